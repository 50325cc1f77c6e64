/* oracle/shim — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Stand-in for google/cpu_features' cpu_features_macros.h, which the reference fetches at
 * configure time (cmake/cpu_features.cmake:4-9) and which is not available offline. */
#pragma once
#if defined(__x86_64__) || defined(_M_X64)
#define CPU_FEATURES_ARCH_X86_64 1
#else
#error "oracle/_ref is only built on x86-64 hosts"
#endif
