// Exact distance kernels: every (row, query) score is computed with the same floating-point
// operation order as the reference's x86 AVX512 dispatch tier, so scores are bit-identical to the
// CPU path (reference: spaces/IP/IP_AVX512F_FP32.h:19-56, L2/L2_AVX512F_FP32.h:21-59,
// IP/IP_AVX512F_FP64.h, IP/IP_AVX512_BF16_VL_BF16.h:23-47, L2/L2_AVX512BW_VBMI2_BF16.h:42-78,
// IP/IP_AVX512F_FP16.h:27-68, IP/IP.cpp:185-286, L2/L2.cpp:76-174, VNNI int8/uint8 kernels).
//
// Mapping: one x86 SIMD lane of one accumulator register == one CUDA thread ("chain"): G threads
// own the G sequential FMA recurrences of a (row, query) pair and finish with the same butterfly
// the CPU's _mm512_reduce_add_ps performs. Rows stream from HBM once per query chunk, R rows per
// thread group are register-tiled against QC queries held in shared memory.
#include "vsgpu_dist.cuh"
#include <cstdlib>

namespace vsgpu {

ChainPlan make_plan(int type, int metric, size_t dim_) {
    ChainPlan p{};
    const int dim = (int)dim_;
    p.dim = dim;
    p.is_l2 = metric == VSGPU_L2;
    auto lanes = [&](int G) {
        p.kind = CK_LANES;
        p.G = G;
        const int L = G / 2;
        p.r = dim % G;
        p.head = p.r % L;
        p.nfull = p.r / L;
        p.prefix = p.r ? 1 : 0;
        p.S = p.prefix + dim / G;
    };
    auto seq = [&]() {
        p.kind = CK_SEQ;
        p.G = 1;
        p.S = dim;
    };
    switch (type) {
    case VSGPU_FLOAT32:
        if (dim < 8) seq(); else lanes(32);
        break;
    case VSGPU_FLOAT64:
        if (dim < 4) seq(); else lanes(16);
        break;
    case VSGPU_BFLOAT16:
        if (dim < 32) {
            seq();
        } else if (p.is_l2) {
            p.kind = CK_BF16_VBMI2;
            p.G = 16;
            p.r = dim % 32;
            p.prefix = (p.r >= 16 ? 1 : 0) + ((p.r % 16) ? 1 : 0);
            p.S = p.prefix + 2 * (dim / 32);
        } else {
            p.kind = CK_BF16_DP;
            p.G = 16;
            p.r = dim % 32;
            p.prefix = p.r ? 1 : 0;
            p.S = 2 * (p.prefix + dim / 32);
            p.ftz = 1;
        }
        break;
    case VSGPU_FLOAT16:
        if (dim < 8) {
            seq();
        } else if (dim < 16) {
            seq();
            p.seq_f16c = 1;
        } else {
            lanes(32);
        }
        break;
    default:
        p.kind = CK_INT;
        p.G = 8;
        p.S = 0;
        break;
    }
    return p;
}

// ------------------------------------------------------------------------------------------------
// queries -> chain layout: qc[q][s][c] (compute type, zero in padded slots)
template <typename CT>
__global__ void prep_queries_kernel(const uint8_t *__restrict__ q, size_t q_stride, int type, ChainPlan plan,
                                    CT *__restrict__ qc, size_t nq) {
    const size_t per_q = (size_t)plan.S * plan.G;
    const size_t total = per_q * nq;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t qi = i / per_q;
        const int rem = (int)(i % per_q);
        const int s = rem / plan.G, c = rem % plan.G;
        const int e = chain_elem(plan, c, s);
        qc[i] = e < 0 ? CT(0) : Loader<CT>::load(q + qi * q_stride, type, e);
    }
}

struct ScanArgs {
    const uint8_t *rows;
    size_t row_stride;
    size_t n;
    int type;
    ChainPlan plan;
    const void *qchain; // [nq][S][G]
    size_t nq;
    void *scores; // [nq][ld]
    size_t ld;
    int q_in_smem;
};

// R rows x QC queries per thread group of G chains.
template <typename CT, int G, bool FTZ, bool L2, int QC, int R>
__global__ void __launch_bounds__(256) exact_scan_kernel(ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int SC = 8;
    constexpr int GROUPS = 32 / G;
    const ChainPlan plan = a.plan;
    const int S = plan.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = lane % G, grp = lane / G;
    const int warps = blockDim.x >> 5;
    const size_t per_q = (size_t)S * G;
    const bool fast = plan.kind == CK_LANES && plan.prefix == 0;

    for (size_t q0 = 0; q0 < a.nq; q0 += QC) {
        const int nqc = (int)min((size_t)QC, a.nq - q0);
        const CT *qbase = reinterpret_cast<const CT *>(a.qchain) + q0 * per_q;
        if (a.q_in_smem) {
            CT *qs = reinterpret_cast<CT *>(smem_raw);
            __syncthreads();
            for (size_t i = threadIdx.x; i < per_q * nqc; i += blockDim.x) qs[i] = qbase[i];
            __syncthreads();
            qbase = qs;
        }
        const size_t rows_per_block = (size_t)warps * GROUPS * R;
        for (size_t tile = (size_t)blockIdx.x * rows_per_block; tile < a.n; tile += (size_t)gridDim.x * rows_per_block) {
            const size_t row0 = tile + ((size_t)warp * GROUPS + grp) * R;
            const uint8_t *rp[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                size_t row = row0 + r;
                if (row >= a.n) row = a.n - 1;
                rp[r] = a.rows + row * a.row_stride;
            }
            CT acc[R][QC];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int q = 0; q < QC; q++) acc[r][q] = CT(0);

            for (int s0 = 0; s0 < S; s0 += SC) {
                CT rv[R][SC];
#pragma unroll
                for (int j = 0; j < SC; j++) {
                    const int s = s0 + j;
                    int e = -1;
                    if (s < S) e = fast ? G * s + c : chain_elem(plan, c, s);
#pragma unroll
                    for (int r = 0; r < R; r++) rv[r][j] = e < 0 ? CT(0) : Loader<CT>::load(rp[r], a.type, e);
                }
#pragma unroll
                for (int q = 0; q < QC; q++) {
                    if (q < nqc) {
                        const CT *qp = qbase + (size_t)q * per_q + (size_t)s0 * G + c;
#pragma unroll
                        for (int j = 0; j < SC; j++) {
                            if (s0 + j < S) {
                                const CT qv = qp[j * G];
#pragma unroll
                                for (int r = 0; r < R; r++) {
                                    if constexpr (L2) {
                                        const CT d = sub_rn(rv[r][j], qv);
                                        acc[r][q] = fma_step<FTZ>(d, d, acc[r][q]);
                                    } else {
                                        acc[r][q] = fma_step<FTZ>(rv[r][j], qv, acc[r][q]);
                                    }
                                }
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < QC; q++) {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    CT v = butterfly<CT, G>(acc[r][q]);
                    if (!L2) v = sub_rn(CT(1), v);
                    if (c == 0 && q < nqc && row0 + r < a.n)
                        reinterpret_cast<CT *>(a.scores)[(q0 + q) * a.ld + row0 + r] = v;
                }
            }
        }
    }
}

// Scalar tiers: the reference's naive loops (multiply and add rounded separately), and the F16C
// fp16 tier for 8 <= dim < 16 (8 lanes, acc0 = head elements, acc1 = next 8, sequential lane sum).
template <typename CT>
__global__ void seq_scan_kernel(const uint8_t *__restrict__ rows, size_t row_stride, size_t n, int type,
                                ChainPlan plan, const uint8_t *__restrict__ q, size_t q_stride, size_t nq,
                                const uint32_t *__restrict__ ids, size_t ids_ld, CT *__restrict__ out, size_t ld,
                                size_t per_q) {
    const size_t total = per_q * nq;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t qi = i / per_q, j = i % per_q;
        size_t row = j;
        if (ids) {
            const uint32_t id = ids[qi * ids_ld + j];
            if (id == 0xffffffffu) continue;
            row = id;
        }
        const uint8_t *rp = rows + row * row_stride;
        const uint8_t *qp = q + qi * q_stride;
        const CT res = seq_dist<CT>(rp, type, plan, [&](int e) { return Loader<CT>::load(qp, type, e); });
        out[qi * ld + j] = res;
    }
}

// Chosen rows (re-rank / ad-hoc distances): one thread group per (query, candidate) pair. Grid = (chunks of
// GATHER_CHUNK candidates, queries): a block reads its query's count once and the chunks past it exit at once (the
// re-rank passes RUN_CAP-wide lists of which a quarter is in use). The chain is sequential in s but the loads are not:
// a group keeps GP pairs x GU steps of row loads in flight and shares the query operand between its pairs (one load
// per FMA on one pair at a time had left the re-rank of a 1024-query batch latency-bound at ~1 TB/s).
constexpr int GATHER_CHUNK = 256;
template <typename CT, int G, bool FTZ, bool L2>
__global__ void __launch_bounds__(256) exact_gather_kernel(ScanArgs a, const uint32_t *__restrict__ ids, size_t ids_ld,
                                                           const uint32_t *__restrict__ counts, size_t max_count) {
    constexpr int GROUPS = 32 / G;
    constexpr int GP = 2, GU = 8;
    const ChainPlan plan = a.plan;
    const int S = plan.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int c = lane % G, grp = lane / G;
    const size_t per_q = (size_t)S * G;
    for (size_t qi = blockIdx.y; qi < a.nq; qi += gridDim.y) {
        size_t cnt = max_count;
        if (counts) cnt = min((size_t)counts[qi], max_count);
        const size_t chunk_end = min(cnt, ((size_t)blockIdx.x + 1) * GATHER_CHUNK);
        const CT *qp = reinterpret_cast<const CT *>(a.qchain) + qi * per_q + c;
        const uint32_t *idq = ids + qi * ids_ld;
        // all groups of a warp iterate the same number of times (the butterfly needs the full warp)
        for (size_t base = (size_t)blockIdx.x * GATHER_CHUNK + (size_t)warp * (GROUPS * GP); base < chunk_end;
             base += (size_t)warps * (GROUPS * GP)) {
            size_t j[GP];
            bool valid[GP];
            const uint8_t *rp[GP];
            CT acc[GP];
#pragma unroll
            for (int p = 0; p < GP; p++) {
                j[p] = base + (size_t)p * GROUPS + grp;
                uint32_t id = 0xffffffffu;
                if (j[p] < chunk_end) id = idq[j[p]];
                valid[p] = id != 0xffffffffu;
                rp[p] = a.rows + (valid[p] ? (size_t)id : 0) * a.row_stride;
                acc[p] = CT(0);
            }
            for (int s0 = 0; s0 < S; s0 += GU) {
                CT x[GP][GU], y[GU];
#pragma unroll
                for (int u = 0; u < GU; u++) {
                    const int s = s0 + u;
                    y[u] = CT(0);
#pragma unroll
                    for (int p = 0; p < GP; p++) x[p][u] = CT(0);
                    if (s < S) {
                        const int e = chain_elem(plan, c, s);
                        y[u] = qp[(size_t)s * G];
                        if (e >= 0) {
#pragma unroll
                            for (int p = 0; p < GP; p++)
                                if (valid[p]) x[p][u] = Loader<CT>::load(rp[p], a.type, e);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < GU; u++) {
                    if (s0 + u < S) {
#pragma unroll
                        for (int p = 0; p < GP; p++) {
                            if constexpr (L2) {
                                const CT d = sub_rn(x[p][u], y[u]);
                                acc[p] = fma_step<FTZ>(d, d, acc[p]);
                            } else {
                                acc[p] = fma_step<FTZ>(x[p][u], y[u], acc[p]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < GP; p++) {
                CT v = butterfly<CT, G>(acc[p]);
                if (!L2) v = sub_rn(CT(1), v);
                if (valid[p] && c == 0) reinterpret_cast<CT *>(a.scores)[qi * a.ld + j[p]] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// integer types: exact in any order. 8 threads per row, 16-byte chunks, dp4a.
struct IntArgs {
    const uint8_t *rows;
    size_t row_stride; // multiple of 16, zero padded
    size_t n;
    int is_unsigned;
    int metric;
    int dim;
    const uint8_t *q; // [nq][row_stride], zero padded
    size_t nq;
    const float *row_norms; // cosine
    const float *q_norms;
    float *scores;
    size_t ld;
    const uint32_t *ids; // gather mode (else NULL)
    size_t ids_ld;
    const uint32_t *counts;
    size_t max_count;
};

template <bool U, int QC>
__global__ void __launch_bounds__(256) int_scan_kernel(IntArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int G = 8, R = 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int c = lane % G, grp = lane / G;
    const int chunks = (int)(a.row_stride / 16);
    uint4 *qs = reinterpret_cast<uint4 *>(smem_raw);
    long long *qq_s = reinterpret_cast<long long *>(smem_raw + (size_t)QC * a.row_stride);
    for (size_t q0 = 0; q0 < a.nq; q0 += QC) {
        const int nqc = (int)min((size_t)QC, a.nq - q0);
        __syncthreads();
        for (int i = threadIdx.x; i < nqc * chunks; i += blockDim.x)
            qs[i] = reinterpret_cast<const uint4 *>(a.q + q0 * a.row_stride)[i];
        __syncthreads();
        if (threadIdx.x < nqc) {
            long long t = 0;
            for (int i = 0; i < chunks; i++) {
                const uint4 v = qs[threadIdx.x * chunks + i];
                t += dot4<U>(v.x, v.x, 0) + (long long)dot4<U>(v.y, v.y, 0) + dot4<U>(v.z, v.z, 0) + (long long)dot4<U>(v.w, v.w, 0);
            }
            qq_s[threadIdx.x] = t;
        }
        __syncthreads();
        const size_t rows_per_block = (size_t)warps * 4 * R;
        for (size_t tile = (size_t)blockIdx.x * rows_per_block; tile < a.n; tile += (size_t)gridDim.x * rows_per_block) {
            const size_t row0 = tile + ((size_t)warp * 4 + grp) * R;
            const uint4 *rp[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                size_t row = row0 + r;
                if (row >= a.n) row = a.n - 1;
                rp[r] = reinterpret_cast<const uint4 *>(a.rows + row * a.row_stride);
            }
            long long dot[R][QC], aa[R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                aa[r] = 0;
#pragma unroll
                for (int q = 0; q < QC; q++) dot[r][q] = 0;
            }
            // int32 partial sums are flushed to 64 bits every 1024 chunks (16 KB): 255^2*16384 < 2^31
            for (int base = 0; base < chunks; base += 1024 * G) {
                int d32[R][QC], a32[R];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    a32[r] = 0;
#pragma unroll
                    for (int q = 0; q < QC; q++) d32[r][q] = 0;
                }
                const int end = min(chunks, base + 1024 * G);
                for (int ch = base + c; ch < end; ch += G) {
                    uint4 rv[R];
#pragma unroll
                    for (int r = 0; r < R; r++) {
                        rv[r] = __ldg(rp[r] + ch);
                        a32[r] = dot4<U>(rv[r].x, rv[r].x, a32[r]);
                        a32[r] = dot4<U>(rv[r].y, rv[r].y, a32[r]);
                        a32[r] = dot4<U>(rv[r].z, rv[r].z, a32[r]);
                        a32[r] = dot4<U>(rv[r].w, rv[r].w, a32[r]);
                    }
#pragma unroll
                    for (int q = 0; q < QC; q++) {
                        if (q < nqc) {
                            const uint4 qv = qs[q * chunks + ch];
#pragma unroll
                            for (int r = 0; r < R; r++) {
                                d32[r][q] = dot4<U>(rv[r].x, qv.x, d32[r][q]);
                                d32[r][q] = dot4<U>(rv[r].y, qv.y, d32[r][q]);
                                d32[r][q] = dot4<U>(rv[r].z, qv.z, d32[r][q]);
                                d32[r][q] = dot4<U>(rv[r].w, qv.w, d32[r][q]);
                            }
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < R; r++) {
                    aa[r] += a32[r];
#pragma unroll
                    for (int q = 0; q < QC; q++) dot[r][q] += d32[r][q];
                }
            }
#pragma unroll
            for (int r = 0; r < R; r++) {
#pragma unroll
                for (int w = G / 2; w >= 1; w >>= 1) aa[r] += __shfl_xor_sync(0xffffffffu, aa[r], w);
#pragma unroll
                for (int q = 0; q < QC; q++) {
#pragma unroll
                    for (int w = G / 2; w >= 1; w >>= 1) dot[r][q] += __shfl_xor_sync(0xffffffffu, dot[r][q], w);
                    if (c == 0 && q < nqc && row0 + r < a.n) {
                        const size_t row = row0 + r;
                        const float rn = a.row_norms ? a.row_norms[row] : 0.f;
                        const float qn = a.q_norms ? a.q_norms[q0 + q] : 0.f;
                        a.scores[(q0 + q) * a.ld + row] = int_score(a.metric, dot[r][q], aa[r], qq_s[q], rn, qn);
                    }
                }
            }
        }
    }
}

template <bool U>
__global__ void __launch_bounds__(256) int_gather_kernel(IntArgs a) {
    constexpr int G = 8;
    const int lane = threadIdx.x & 31;
    const int c = lane % G, grp = lane / G;
    const int chunks = (int)(a.row_stride / 16);
    const size_t group_id = ((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + grp;
    const size_t n_groups = (size_t)gridDim.x * (blockDim.x >> 5) * 4;
    const size_t total = a.nq * a.max_count;
    const size_t iters = (total + n_groups - 1) / n_groups;
    for (size_t it = 0; it < iters; it++) {
        const size_t pair = it * n_groups + group_id;
        bool valid = pair < total;
        size_t qi = 0, j = 0;
        uint32_t id = 0xffffffffu;
        if (valid) {
            qi = pair / a.max_count;
            j = pair % a.max_count;
            const uint32_t cnt = a.counts ? a.counts[qi] : (uint32_t)a.max_count;
            if (j < cnt) id = a.ids[qi * a.ids_ld + j];
            valid = id != 0xffffffffu;
        }
        const uint4 *rp = reinterpret_cast<const uint4 *>(a.rows + (valid ? (size_t)id : 0) * a.row_stride);
        const uint4 *qp = reinterpret_cast<const uint4 *>(a.q + qi * a.row_stride);
        long long dot = 0, aa = 0, qq = 0;
        if (valid) {
            for (int ch = c; ch < chunks; ch += G) {
                const uint4 rv = __ldg(rp + ch), qv = __ldg(qp + ch);
                dot += (long long)dot4<U>(rv.x, qv.x, 0) + dot4<U>(rv.y, qv.y, 0) + (long long)dot4<U>(rv.z, qv.z, 0) + dot4<U>(rv.w, qv.w, 0);
                aa += (long long)dot4<U>(rv.x, rv.x, 0) + dot4<U>(rv.y, rv.y, 0) + (long long)dot4<U>(rv.z, rv.z, 0) + dot4<U>(rv.w, rv.w, 0);
                qq += (long long)dot4<U>(qv.x, qv.x, 0) + dot4<U>(qv.y, qv.y, 0) + (long long)dot4<U>(qv.z, qv.z, 0) + dot4<U>(qv.w, qv.w, 0);
            }
        }
#pragma unroll
        for (int w = G / 2; w >= 1; w >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, w);
            aa += __shfl_xor_sync(0xffffffffu, aa, w);
            qq += __shfl_xor_sync(0xffffffffu, qq, w);
        }
        if (valid && c == 0) {
            const float rn = a.row_norms ? a.row_norms[id] : 0.f;
            const float qn = a.q_norms ? a.q_norms[qi] : 0.f;
            a.scores[qi * a.ld + j] = int_score(a.metric, dot, aa, qq, rn, qn);
        }
    }
}

// ------------------------------------------------------------------------------------------------
static int grid_for(int device, int blocks_per_sm) {
    static int sms[64] = {0};
    if (device < 0 || device >= 64) return 148 * blocks_per_sm;
    if (!sms[device]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
        sms[device] = v;
    }
    return sms[device] * blocks_per_sm;
}

template <typename CT, int G, bool FTZ, bool L2>
static int launch_scan_t(vsgpu_store *s, ScanArgs &a) {
    const size_t per_q_bytes = (size_t)a.plan.S * G * sizeof(CT);
    constexpr int R = sizeof(CT) == 8 ? 2 : 4;
    // query chunk: as many as fit in ~96 KB of shared memory, capped at 16 (accumulator registers)
    int qc = 16;
    if (a.nq <= 1) qc = 1;
    else if (a.nq <= 4) qc = 4;
    size_t smem = per_q_bytes * qc;
    const size_t smem_cap = 160 * 1024;
    while (qc > 1 && smem > smem_cap) {
        qc = qc == 16 ? 4 : 1;
        smem = per_q_bytes * qc;
    }
    a.q_in_smem = smem <= smem_cap;
    if (!a.q_in_smem) smem = 0;
    const size_t rows_per_block = 8 * (32 / G) * R;
    size_t blocks = (a.n + rows_per_block - 1) / rows_per_block;
    const size_t max_blocks = (size_t)grid_for(s->device, 2);
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks == 0) blocks = 1;
#define VS_LAUNCH_SCAN(QCV)                                                                                    \
    do {                                                                                                       \
        auto kern = exact_scan_kernel<CT, G, FTZ, L2, QCV, R>;                                                 \
        if (smem > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<(unsigned)blocks, 256, smem, s->stream>>>(a);                                                   \
    } while (0)
    if (qc == 16) VS_LAUNCH_SCAN(16);
    else if (qc == 4) VS_LAUNCH_SCAN(4);
    else VS_LAUNCH_SCAN(1);
#undef VS_LAUNCH_SCAN
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

template <typename CT, int G, bool FTZ, bool L2>
static int launch_gather_t(vsgpu_store *s, ScanArgs &a, const uint32_t *ids, size_t ids_ld, const uint32_t *counts,
                           size_t max_count) {
    const dim3 blocks((unsigned)((max_count + GATHER_CHUNK - 1) / GATHER_CHUNK), (unsigned)std::min<size_t>(a.nq, 65535));
    exact_gather_kernel<CT, G, FTZ, L2><<<blocks, 256, 0, s->stream>>>(a, ids, ids_ld, counts, max_count);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

// chain-layout copy of the queries lives in s->misc
static int prep_queries(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const void **qchain) {
    const ChainPlan &p = s->plan;
    const size_t csz = s->type == VSGPU_FLOAT64 ? 8 : 4;
    const size_t bytes = (size_t)p.S * p.G * nq * csz;
    VS_TRY(ensure_scratch(s, s->misc, bytes));
    const size_t total = (size_t)p.S * p.G * nq;
    unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 4096);
    if (blocks == 0) blocks = 1;
    if (csz == 8)
        prep_queries_kernel<double><<<blocks, 256, 0, s->stream>>>((const uint8_t *)q_dev, q_stride, s->type, p, (double *)s->misc.ptr, nq);
    else
        prep_queries_kernel<float><<<blocks, 256, 0, s->stream>>>((const uint8_t *)q_dev, q_stride, s->type, p, (float *)s->misc.ptr, nq);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    *qchain = s->misc.ptr;
    return VSGPU_OK;
}

static int launch_int(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms, void *scores,
                      size_t ld, const uint32_t *ids, size_t ids_ld, const uint32_t *counts, size_t max_count) {
    if (q_stride != s->row_stride) {
        set_error("integer queries must be staged with the store's padded row stride");
        return VSGPU_ERR_ARG;
    }
    IntArgs a{};
    a.rows = s->rows;
    a.row_stride = s->row_stride;
    a.n = s->count;
    a.is_unsigned = s->type == VSGPU_UINT8;
    a.metric = s->metric;
    a.dim = (int)s->dim;
    a.q = (const uint8_t *)q_dev;
    a.nq = nq;
    a.row_norms = s->has_norm ? s->norms : nullptr;
    a.q_norms = s->has_norm ? q_norms : nullptr;
    a.scores = (float *)scores;
    a.ld = ld;
    a.ids = ids;
    a.ids_ld = ids_ld;
    a.counts = counts;
    a.max_count = max_count;
    if (ids) {
        const size_t pairs = nq * max_count;
        size_t blocks = std::min<size_t>((pairs + 31) / 32, (size_t)grid_for(s->device, 8));
        if (blocks == 0) blocks = 1;
        if (a.is_unsigned) int_gather_kernel<true><<<(unsigned)blocks, 256, 0, s->stream>>>(a);
        else int_gather_kernel<false><<<(unsigned)blocks, 256, 0, s->stream>>>(a);
    } else {
        int qc = nq <= 1 ? 1 : (nq <= 4 ? 4 : 16);
        size_t smem = (size_t)qc * s->row_stride + 16 * sizeof(long long);
        while (qc > 1 && smem > 160 * 1024) {
            qc = qc == 16 ? 4 : 1;
            smem = (size_t)qc * s->row_stride + 16 * sizeof(long long);
        }
        if (smem > 200 * 1024) {
            set_error("dimension too large for the integer scan kernel");
            return VSGPU_ERR_ARG;
        }
        const size_t rows_per_block = 8 * 4 * 2;
        size_t blocks = std::min<size_t>((a.n + rows_per_block - 1) / rows_per_block, (size_t)grid_for(s->device, 2));
        if (blocks == 0) blocks = 1;
#define VS_LAUNCH_INT(UV, QCV)                                                                                 \
    do {                                                                                                       \
        auto kern = int_scan_kernel<UV, QCV>;                                                                  \
        if (smem > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<(unsigned)blocks, 256, smem, s->stream>>>(a);                                                   \
    } while (0)
        if (a.is_unsigned) {
            if (qc == 16) VS_LAUNCH_INT(true, 16); else if (qc == 4) VS_LAUNCH_INT(true, 4); else VS_LAUNCH_INT(true, 1);
        } else {
            if (qc == 16) VS_LAUNCH_INT(false, 16); else if (qc == 4) VS_LAUNCH_INT(false, 4); else VS_LAUNCH_INT(false, 1);
        }
#undef VS_LAUNCH_INT
    }
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

static int launch_seq(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, void *scores, size_t ld,
                      const uint32_t *ids, size_t ids_ld, size_t per_q) {
    const size_t total = per_q * nq;
    unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, (size_t)grid_for(s->device, 8));
    if (blocks == 0) blocks = 1;
    if (s->type == VSGPU_FLOAT64)
        seq_scan_kernel<double><<<blocks, 256, 0, s->stream>>>(s->rows, s->row_stride, s->count, s->type, s->plan,
                                                               (const uint8_t *)q_dev, q_stride, nq, ids, ids_ld,
                                                               (double *)scores, ld, per_q);
    else
        seq_scan_kernel<float><<<blocks, 256, 0, s->stream>>>(s->rows, s->row_stride, s->count, s->type, s->plan,
                                                              (const uint8_t *)q_dev, q_stride, nq, ids, ids_ld,
                                                              (float *)scores, ld, per_q);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

template <typename F> static int dispatch_chain(const vsgpu_store *s, F &&f) {
    const ChainPlan &p = s->plan;
    const bool l2 = p.is_l2;
    if (s->type == VSGPU_FLOAT64) return l2 ? f.template operator()<double, 16, false, true>() : f.template operator()<double, 16, false, false>();
    if (p.kind == CK_BF16_DP) return f.template operator()<float, 16, true, false>();
    if (p.kind == CK_BF16_VBMI2) return f.template operator()<float, 16, false, true>();
    return l2 ? f.template operator()<float, 32, false, true>() : f.template operator()<float, 32, false, false>();
}

int launch_exact_scan(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms, void *scores,
                      size_t ld) {
    if (s->count == 0 || nq == 0) return VSGPU_OK;
    const ChainPlan &p = s->plan;
    if (p.kind == CK_INT) return launch_int(s, q_dev, nq, q_stride, q_norms, scores, ld, nullptr, 0, nullptr, 0);
    if (p.kind == CK_SEQ) return launch_seq(s, q_dev, nq, q_stride, scores, ld, nullptr, 0, s->count);
    static const bool legacy = getenv("VSGPU_LEGACY_SCAN") != nullptr; // A/B switch for profiling
    if (!legacy && nq <= 16 && tma_scan_supported(s)) return launch_tma_scan(s, q_dev, nq, q_stride, scores, ld);
    const void *qchain = nullptr;
    VS_TRY(prep_queries(s, q_dev, nq, q_stride, &qchain));
    ScanArgs a{};
    a.rows = s->rows;
    a.row_stride = s->row_stride;
    a.n = s->count;
    a.type = s->type;
    a.plan = p;
    a.qchain = qchain;
    a.nq = nq;
    a.scores = scores;
    a.ld = ld;
    return dispatch_chain(s, [&]<typename CT, int G, bool FTZ, bool L2>() { return launch_scan_t<CT, G, FTZ, L2>(s, a); });
}

int launch_exact_gather(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms,
                        const uint32_t *ids, size_t ids_ld, const uint32_t *counts, size_t max_count, void *out,
                        size_t ld) {
    if (nq == 0 || max_count == 0) return VSGPU_OK;
    const ChainPlan &p = s->plan;
    if (p.kind == CK_INT) return launch_int(s, q_dev, nq, q_stride, q_norms, out, ld, ids, ids_ld, counts, max_count);
    if (p.kind == CK_SEQ) {
        if (counts) {
            set_error("sequential tier gather does not take per-query counts");
            return VSGPU_ERR_ARG;
        }
        return launch_seq(s, q_dev, nq, q_stride, out, ld, ids, ids_ld, max_count);
    }
    const void *qchain = nullptr;
    VS_TRY(prep_queries(s, q_dev, nq, q_stride, &qchain));
    ScanArgs a{};
    a.rows = s->rows;
    a.row_stride = s->row_stride;
    a.n = s->count;
    a.type = s->type;
    a.plan = p;
    a.qchain = qchain;
    a.nq = nq;
    a.scores = out;
    a.ld = ld;
    return dispatch_chain(s, [&]<typename CT, int G, bool FTZ, bool L2>() {
        return launch_gather_t<CT, G, FTZ, L2>(s, a, ids, ids_ld, counts, max_count);
    });
}

} // namespace vsgpu
