/* oracle/shim — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * Minimal replacement for google/cpu_features' cpuinfo_x86.h: just the feature bits that
 * the reference's dispatchers read (`grep -o 'features\.[a-z0-9_]*' src/VecSim/spaces/*.cpp`).
 * Bits come straight from CPUID/XGETBV. `vsref_feature_disable_mask` lets the test harness
 * walk the dispatcher down tier by tier, the way tests/unit/test_spaces.cpp:703-709 does by
 * zeroing fields of a features copy. */
#pragma once
#include <cpuid.h>
#include <stdint.h>

extern "C" uint32_t vsref_feature_disable_mask; /* defined in ref_harness.cpp */

namespace cpu_features {

struct X86Features {
    int sse, sse3, sse4_1, avx, avx2, fma3, f16c;
    int avx512f, avx512bw, avx512vl, avx512vnni, avx512vbmi2, avx512_bf16, avx512_fp16;
};
struct X86Info {
    X86Features features;
};

enum VsrefFeatureBit {
    VSREF_SSE = 1u << 0, VSREF_SSE3 = 1u << 1, VSREF_SSE4_1 = 1u << 2, VSREF_AVX = 1u << 3,
    VSREF_AVX2 = 1u << 4, VSREF_FMA3 = 1u << 5, VSREF_F16C = 1u << 6, VSREF_AVX512F = 1u << 7,
    VSREF_AVX512BW = 1u << 8, VSREF_AVX512VL = 1u << 9, VSREF_AVX512VNNI = 1u << 10,
    VSREF_AVX512VBMI2 = 1u << 11, VSREF_AVX512_BF16 = 1u << 12, VSREF_AVX512_FP16 = 1u << 13,
};

static inline uint64_t vsref_xgetbv0() {
    uint32_t lo, hi;
    __asm__ volatile("xgetbv" : "=a"(lo), "=d"(hi) : "c"(0));
    return ((uint64_t)hi << 32) | lo;
}

static inline X86Info GetX86Info() {
    X86Info info{};
    unsigned a, b, c, d;
    unsigned max_leaf = __get_cpuid_max(0, nullptr);
    __cpuid(1, a, b, c, d);
    const bool osxsave = (c >> 27) & 1;
    uint64_t xcr0 = osxsave ? vsref_xgetbv0() : 0;
    const bool os_avx = (xcr0 & 0x6) == 0x6;
    const bool os_avx512 = os_avx && (xcr0 & 0xe0) == 0xe0;
    X86Features &f = info.features;
    f.sse = (d >> 25) & 1;
    f.sse3 = c & 1;
    f.sse4_1 = (c >> 19) & 1;
    f.avx = os_avx && ((c >> 28) & 1);
    f.fma3 = os_avx && ((c >> 12) & 1);
    f.f16c = os_avx && ((c >> 29) & 1);
    if (max_leaf >= 7) {
        __cpuid_count(7, 0, a, b, c, d);
        f.avx2 = os_avx && ((b >> 5) & 1);
        f.avx512f = os_avx512 && ((b >> 16) & 1);
        f.avx512bw = os_avx512 && ((b >> 30) & 1);
        f.avx512vl = os_avx512 && ((b >> 31) & 1);
        f.avx512vbmi2 = os_avx512 && ((c >> 6) & 1);
        f.avx512vnni = os_avx512 && ((c >> 11) & 1);
        f.avx512_fp16 = os_avx512 && ((d >> 23) & 1);
        unsigned a1, b1, c1, d1;
        __cpuid_count(7, 1, a1, b1, c1, d1);
        f.avx512_bf16 = os_avx512 && ((a1 >> 5) & 1);
    }
    const uint32_t off = vsref_feature_disable_mask;
    if (off & VSREF_SSE) f.sse = 0;
    if (off & VSREF_SSE3) f.sse3 = 0;
    if (off & VSREF_SSE4_1) f.sse4_1 = 0;
    if (off & VSREF_AVX) f.avx = 0;
    if (off & VSREF_AVX2) f.avx2 = 0;
    if (off & VSREF_FMA3) f.fma3 = 0;
    if (off & VSREF_F16C) f.f16c = 0;
    if (off & VSREF_AVX512F) f.avx512f = 0;
    if (off & VSREF_AVX512BW) f.avx512bw = 0;
    if (off & VSREF_AVX512VL) f.avx512vl = 0;
    if (off & VSREF_AVX512VNNI) f.avx512vnni = 0;
    if (off & VSREF_AVX512VBMI2) f.avx512vbmi2 = 0;
    if (off & VSREF_AVX512_BF16) f.avx512_bf16 = 0;
    if (off & VSREF_AVX512_FP16) f.avx512_fp16 = 0;
    return info;
}

} // namespace cpu_features
