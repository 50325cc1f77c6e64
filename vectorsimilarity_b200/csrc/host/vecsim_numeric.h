// Host-side numeric helpers of the product path: type conversions and the cosine preprocessing that
// the reference applies once per stored vector / query (not hot, but part of result parity).
// Behaviour restated from the reference (paths relative to /root/reference/src/VecSim):
//   types/bfloat16.h:23-39, types/float16.h:33-117, spaces/normalize/normalize_naive.h:23-88,
//   spaces/normalize/compute_norm.h:17-31. Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace vsb {

inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// fp32 -> bf16, round-to-nearest-even on the dropped half
inline uint16_t f32_to_bf16(float f) {
    uint32_t u = f2u(f);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
inline float bf16_to_f32(uint16_t h) { return u2f((uint32_t)h << 16); }

// fp16 -> fp32, exact
inline float fp16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t mag = h & 0x7fffu;
    uint32_t bits;
    if ((mag & 0x7c00u) == 0x7c00u) {
        bits = 0x7f800000u | ((mag & 0x3ffu) << 13);                 // inf / nan
    } else if ((mag & 0x7c00u) == 0) {
        // subnormal half = mag * 2^-24, exactly representable
        bits = f2u((float)mag * u2f(0x33800000u));
    } else {
        bits = (mag << 13) + (112u << 23);
    }
    return u2f(bits | sign);
}
// fp32 -> fp16 as the reference does it: the low 12 mantissa bits are dropped first, then the value
// is rounded half-up at bit 12 (float16.h:60-117); NOT IEEE round-to-nearest-even.
inline uint16_t f32_to_fp16(float f) {
    uint32_t mag = f2u(f);
    const uint32_t sign = mag & 0x80000000u;
    mag ^= sign;
    const uint32_t inf = 255u << 23;
    uint32_t out = mag > inf ? 0x7e00u : 0x7c00u;
    if (mag < inf) {
        float scaled = u2f(mag & ~0xfffu) * u2f(15u << 23);
        const float cap = u2f((31u << 23) - 0x1000u);
        if (cap < scaled) scaled = cap;
        const int32_t t = (int32_t)(f2u(scaled) + 0x1000u);
        out = (uint32_t)(t >> 13);
    }
    return (uint16_t)(out | (sign >> 16));
}

inline void normalize_f32(float *v, size_t dim) {
    double sum = 0;
    for (size_t i = 0; i < dim; i++) sum += (double)v[i] * (double)v[i];
    const float norm = (float)std::sqrt(sum);
    for (size_t i = 0; i < dim; i++) v[i] = v[i] / norm;
}
inline void normalize_f64(double *v, size_t dim) {
    double sum = 0;
    for (size_t i = 0; i < dim; i++) sum += v[i] * v[i];
    const double norm = std::sqrt(sum);
    for (size_t i = 0; i < dim; i++) v[i] = v[i] / norm;
}
template <bool BF> inline void normalize_16(uint16_t *v, size_t dim) {
    std::vector<float> tmp(dim);
    float sum = 0;
    for (size_t i = 0; i < dim; i++) {
        tmp[i] = BF ? bf16_to_f32(v[i]) : fp16_to_f32(v[i]);
        sum += tmp[i] * tmp[i];
    }
    const float norm = std::sqrt(sum);
    for (size_t i = 0; i < dim; i++) v[i] = BF ? f32_to_bf16(tmp[i] / norm) : f32_to_fp16(tmp[i] / norm);
}
// int8 / uint8: elements untouched, fp32 norm written behind them
template <typename T> inline void append_int_norm(void *blob, size_t dim) {
    const T *v = (const T *)blob;
    uint64_t sum = 0;
    for (size_t i = 0; i < dim; i++) sum += (uint64_t)((int)v[i] * (int)v[i]);
    const float norm = (float)std::sqrt((double)sum);
    std::memcpy((char *)blob + dim, &norm, sizeof(norm));
}

} // namespace vsb
