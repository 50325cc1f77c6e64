// Device-side distance arithmetic shared by the flat scan (vsgpu_exact.cu) and the HNSW traversal
// (vsgpu_hnsw.cu): exact element conversions, correctly rounded steps, the lane butterfly, the
// integer score formulas. See DESIGN.md §3 for the reference kernels each piece reproduces.
#pragma once
#include "vsgpu_internal.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace vsgpu {

// ------------------------------------------------------------------------------------------------
// element loads: stored type -> compute type, exact conversions only
template <typename CT> struct Loader;
template <> struct Loader<float> {
    static __device__ __forceinline__ float load(const uint8_t *row, int type, int e) {
        if (type == VSGPU_FLOAT32) return __ldg(reinterpret_cast<const float *>(row) + e);
        const unsigned short h = __ldg(reinterpret_cast<const unsigned short *>(row) + e);
        if (type == VSGPU_BFLOAT16) return __uint_as_float((unsigned)h << 16);
        return __half2float(__ushort_as_half(h));
    }
};
template <> struct Loader<double> {
    static __device__ __forceinline__ double load(const uint8_t *row, int, int e) {
        return __ldg(reinterpret_cast<const double *>(row) + e);
    }
};

template <bool FTZ> __device__ __forceinline__ float fma_step(float a, float b, float c) {
    if constexpr (FTZ) {
        float d;
        asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
        return d;
    } else {
        return __fmaf_rn(a, b, c);
    }
}
template <bool FTZ> __device__ __forceinline__ double fma_step(double a, double b, double c) { return __fma_rn(a, b, c); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

template <typename CT, int G> __device__ __forceinline__ CT butterfly(CT v) {
#pragma unroll
    for (int w = G / 2; w >= 1; w >>= 1) v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, w));
    return v;
}


// Scalar tiers: the reference's naive loops (IP/IP.cpp:185-238, L2/L2.cpp:76-133: multiply and add
// rounded separately), and the F16C fp16 tier for 8 <= dim < 16 (IP_F16C_FP16.h:28-80 /
// L2_F16C_FP16.h:28-82: 8 lanes, sum0 lanes < head hold the first `head` elements, sum1 the next 8;
// then lanes 0..7 summed in order). `q(e)` yields element e of the other operand as CT.
template <typename CT, typename QF>
__device__ __forceinline__ CT seq_dist(const uint8_t *rp, int type, const ChainPlan &plan, QF q) {
    CT res;
    if (plan.seq_f16c) {
        const int head = plan.dim % 8;
        CT tot = CT(0);
        for (int l = 0; l < 8; l++) {
            CT s0 = CT(0), s1;
            if (l < head) {
                const CT x = Loader<CT>::load(rp, type, l), y = q(l);
                if (plan.is_l2) { const CT d = sub_rn(x, y); s0 = mul_rn(d, d); } else s0 = mul_rn(x, y);
            }
            const CT x = Loader<CT>::load(rp, type, head + l), y = q(head + l);
            if (plan.is_l2) { const CT d = sub_rn(x, y); s1 = fma_step<false>(d, d, CT(0)); } else s1 = fma_step<false>(x, y, CT(0));
            const CT lane_sum = add_rn(add_rn(s0, s1), CT(0));
            tot = l == 0 ? lane_sum : add_rn(tot, lane_sum);
        }
        res = tot;
    } else {
        res = CT(0);
        for (int e = 0; e < plan.dim; e++) {
            const CT x = Loader<CT>::load(rp, type, e), y = q(e);
            if (plan.is_l2) {
                const CT d = sub_rn(x, y);
                res = add_rn(res, mul_rn(d, d));
            } else {
                res = add_rn(res, mul_rn(x, y));
            }
        }
    }
    if (!plan.is_l2) res = sub_rn(CT(1), res);
    return res;
}

// integer types: exact in any order
template <bool U> __device__ __forceinline__ int dot4(unsigned a, unsigned b, int c) {
    if constexpr (U) return (int)__dp4a(a, b, (unsigned)c);
    else return __dp4a((int)a, (int)b, c);
}

__device__ __forceinline__ float int_score(int metric, long long dot, long long aa, long long qq, float rn, float qn) {
    if (metric == VSGPU_L2) return __ll2float_rn(aa + qq - 2 * dot);     // float(sum (a-b)^2), L2.cpp:164-174
    if (metric == VSGPU_IP) return __ll2float_rn(1 - dot);                // float(1 - sum), IP.cpp:258-277
    // 1.0f - float(ip) / (norm_a * norm_b), IP_AVX512F_BW_VL_VNNI_INT8.h:70-77
    return __fsub_rn(1.0f, __fdiv_rn(__ll2float_rn(dot), __fmul_rn(rn, qn)));
}


} // namespace vsgpu
