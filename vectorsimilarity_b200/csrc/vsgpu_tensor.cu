// Tensor-core coarse pass + exact re-rank for large query batches.
//
// Batched IP / Cosine over contiguous rows is a dense Q x V^T contraction (north star), so for
// nq >= 32 the scan runs on the 5th-gen tensor cores:
//   * rows are the MMA "A" operand (128 rows per tile), streamed HBM -> smem by TMA as 128B-swizzled
//     [128 x 64] bf16 boxes; queries are the "B" operand ([256 x 64] boxes, L2 resident);
//     tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) accumulates a 128 x 256 tile in TMEM, double
//     buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1;
//   * the epilogue reads the accumulators with tcgen05.ld and keeps only rows whose coarse score can
//     still beat the query's current k-th exact score (+ a proven error bound): a handful per tile.
//     Nothing but those candidate ids ever leaves the SM;
//   * candidates are re-scored by the exact kernels (vsgpu_exact.cu: bit-identical to the CPU
//     reference) and merged into the running top-k, which tightens the bound for the next phase.
// fp32 stores keep a bf16 (RNE) mirror of the rows for the coarse pass; bf16 stores are used as is.
// The result is exactly the exact path's result: the bound makes the candidate set a superset of
// the true top-k (proof in DESIGN.md §5.3), and a query whose candidate buffer overflows is redone
// on the exact path.
#include "vsgpu_tc.cuh"
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include <vector>

namespace vsgpu {

// ------------------------------------------------------------------------------------------------
// tile configuration
constexpr int BM = 128;          // rows per tile (UMMA M, cta_group::1)
constexpr int BN = 256;          // queries per tile (UMMA N)
constexpr int BK = 64;           // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_BYTES = BN * BK * 2;  // 32 KB
constexpr int MAX_NQ = 4096;          // thresholds of the whole batch sit in smem
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
// candidate ids per query and phase / survivors kept per query between phases: the small pair serves k <= 384 (the merge
// kernel's tables fit 48 KB of shared memory, seven blocks per SM), the large pair k up to 1024 (docs/benchmarks.md:55-64
// of the reference runs k = 500)
constexpr uint32_t CAND_CAP = 3072, CAND_CAP_BIG = 8192;
constexpr uint32_t RUN_CAP_BIG = 4096;
constexpr size_t K_SMALL_MAX = 384, K_MAX = 1024;

struct GemmSmem {
    // operand ring (1024-byte aligned for SWIZZLE_128B)
    uint8_t a[STAGES][A_BYTES];
    uint8_t b[STAGES][B_BYTES];
    float athr[MAX_NQ];
    float e1[MAX_NQ];
    uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
};

// kind::f16 instruction descriptor: D fp32 (bits 4-5 = 1), A/B bf16 (bits 7-9, 10-12 = 1), both
// K-major, N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t IDESC_BF16 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
// same with A/B format 0 = fp16 (fp16 stores go through the MMA as they are: fp16 x fp16 products are exact in fp32)
constexpr uint32_t IDESC_FP16 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct GemmArgs {
    uint32_t row0;       // first row of this phase (multiple of BM)
    uint32_t row_end;    // one past the last row
    uint32_t nq;         // valid queries
    uint32_t n_qtiles;   // ceil(nq / BN)
    uint32_t k_blocks;   // ceil(dim / BK)
    const float *athr;   // [nq] admit when acc + e1[q] * ||row|| >= athr[q]
    const float *e1;     // [nq] error of the coarse score per unit of row norm
    const float *row_l2; // [rows] ||row||_2, rounded up
    float c_l2;          // L2: the reference's own fp32 score rounds within c_l2 (||q|| + ||row||)^2
    uint32_t *cnt;       // [nq] candidate counters
    uint2 *cand;         // [nq][cand_cap] (row id, coarse accumulator bits)
    uint32_t cand_cap;
    float *dump;         // debug: [rows][dump_ld] raw accumulators (else NULL)
    uint32_t dump_ld;
    uint32_t idesc;      // IDESC_BF16 / IDESC_FP16
    const float *row_sub; // L2: ||row||^2 / 2, subtracted from the accumulator (SUB kernels)
};

// SUB = false: IP / Cosine, the accumulator itself is filtered. SUB = true: L2 — score = ||a||^2 + ||q||^2 - 2 a.q, so with
// ||q||^2 constant per query the rows are ranked by a' = a.q - ||a||^2 / 2 (larger is better); the epilogue subtracts the
// row's half square (one FADD per accumulator) and everything downstream (bounds, merge) works on a'.
template <bool SUB>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
coarse_gemm_filter_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    GemmSmem &sm = *reinterpret_cast<GemmSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m_tiles = (g.row_end - g.row0 + BM - 1) / BM;
    const uint32_t items = m_tiles * g.n_qtiles;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; i++) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tfull[i], 1);
            mbar_init(&sm.tempty[i], EPI_WARPS * 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (uint32_t i = threadIdx.x; i < g.n_qtiles * BN; i += blockDim.x) {
        sm.athr[i] = i < g.nq ? g.athr[i] : __int_as_float(0x7f800000);
        sm.e1[i] = (i < g.nq && g.e1) ? g.e1[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
                const uint32_t mt = item / g.n_qtiles, nt = item % g.n_qtiles;
                const int row = (int)(g.row0 + mt * BM), qrow = (int)(nt * BN);
                for (uint32_t kb = 0; kb < g.k_blocks; kb++) {
                    mbar_wait(&sm.empty[stage], phase ^ 1);
                    mbar_expect_tx(&sm.full[stage], A_BYTES + B_BYTES);
                    tma_load_2d(sm.a[stage], &map_a, &sm.full[stage], (int)(kb * BK), row);
                    tma_load_2d(sm.b[stage], &map_b, &sm.full[stage], (int)(kb * BK), qrow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
                mbar_wait(&sm.tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + as * BN;
                for (uint32_t kb = 0; kb < g.k_blocks; kb++) {
                    mbar_wait(&sm.full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_desc(smem_u32(sm.a[stage]));
                    const uint64_t bdesc = make_desc(smem_u32(sm.b[stage]));
#pragma unroll
                    for (int k = 0; k < BK / UK; k++) {
                        // advance 32 bytes (2 x 16 B) along K inside the swizzled row
                        tc_mma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), g.idesc, (kb | (uint32_t)k) != 0);
                    }
                    tc_commit(&sm.empty[stage]);   // frees the smem slot once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&sm.tfull[as]);          // accumulator complete
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> threshold filter -> candidate ids =====
        const int ew = warp - 4;
        const uint32_t quad = (uint32_t)(warp & 3);       // TMEM lane quadrant this warp may read
        const uint32_t half = (uint32_t)(ew >> 2);         // which 128 of the 256 columns
        uint32_t as = 0, aphase = 0;
        // the per-row terms of the NEXT tile are fetched while this one is filtered (no global load in front of a tile's
        // first compare)
        auto row_terms = [&](uint32_t item, float &nr_o, float &sub_o) {
            const uint32_t r = g.row0 + (item / g.n_qtiles) * BM + quad * 32 + (uint32_t)lane;
            const bool ok = item < items && r < g.row_end;
            nr_o = (ok && g.row_l2) ? __ldg(g.row_l2 + r) : 0.f;
            sub_o = 0.f;
            if constexpr (SUB) sub_o = ok ? __ldg(g.row_sub + r) : 0.f;
        };
        float nr_n, sub_n;
        row_terms(blockIdx.x, nr_n, sub_n);
        DeferredHits dh;
        dh.init();
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const uint32_t mt = item / g.n_qtiles, nt = item % g.n_qtiles;
            const uint32_t row = g.row0 + mt * BM + quad * 32 + (uint32_t)lane;
            const bool row_ok = row < g.row_end;
            const float nr = nr_n, sub = sub_n;
            row_terms(item + gridDim.x, nr_n, sub_n);
            float ra = 0.f;
            // L2: what is filtered (and stored) is a'' = a.q - ||a||^2 / 2 + c_l2 ||a||^2 — the row's own share of the L2
            // rounding term rides on the per-row constant, so it costs nothing per accumulator (rounded so a'' errs upwards)
            if constexpr (SUB) ra = row_ok ? __fmaf_rd(__fmul_ru(g.c_l2, nr), -nr, sub) : 0.f;
            (void)sub;
            mbar_wait(&sm.tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (uint32_t c = 0; c < 4; c++) {
                const uint32_t col = half * 128 + c * 32;
                uint32_t r[32];
                tc_ld32(tmem + ((quad * 32) << 16) + as * BN + col, r);
                tc_wait_ld(r);
                float thr[32];
                lds_f32x32(smem_u32(&sm.athr[nt * BN + col]), thr);
                if (g.dump) {
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const uint32_t q = nt * BN + col + j;
                            if (q < g.nq) g.dump[(size_t)(row - g.row0) * g.dump_ld + q] = __uint_as_float(r[j]);
                        }
                    }
                } else {
                    // admit when the row's coarse score plus ITS error bound e1[q] * ||row|| reaches the query's bound: one FFMA
                    // and one compare per accumulator, all 32 first (independent, no branch in between), then the hits
                    // of the whole warp
                    uint32_t hit = 0;
                    if constexpr (SUB) {
#pragma unroll
                        for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__fsub_rn(__uint_as_float(r[j]), ra));
                    }
                    float e1v[32];
                    lds_f32x32(smem_u32(&sm.e1[nt * BN + col]), e1v);
#pragma unroll
                    for (int j = 0; j < 32; j++) hit |= (__fmaf_rn(e1v[j], nr, __uint_as_float(r[j])) >= thr[j] ? 1u : 0u) << j;
                    warp_append_hits_deferred(row_ok ? hit : 0u, nt * BN + col, row, r, g.cnt, g.cand, lane, g.cand_cap, dh);
                }
            }
            tc_fence_before();
            mbar_arrive(&sm.tempty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        dh.flush(g.cand, g.cand_cap);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (VSGPU_GEMM_PAIR=1): a cluster of two CTAs computes a 256 x 256 tile with tcgen05.mma.cta_group::2.
// Each CTA stages its own 128 rows and HALF of the query tile (128 queries), so a pair moves (256 + 256) operand rows per
// 256 x 256 tile where two single CTAs move 2 x (128 + 256): a third less L2 -> SM traffic, which is what bounds the
// single-CTA kernel. The leader (cluster rank 0) issues the MMAs; operands of both CTAs complete on the
// leader's `full` barriers, commits are multicast to both CTAs' `empty` / `tfull` barriers, and both CTAs' epilogue warps
// release the accumulator on the leader's `tempty`.
constexpr int PSTAGES = 6;
constexpr int BH_BYTES = (BN / 2) * BK * 2; // 16 KB: this CTA's half of the query tile
struct PairSmem {
    uint8_t a[PSTAGES][A_BYTES];
    uint8_t b[PSTAGES][BH_BYTES];
    float athr[MAX_NQ];
    float e1[MAX_NQ];
    uint64_t full[PSTAGES], empty[PSTAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
};
__host__ __device__ constexpr uint32_t idesc_pair(uint32_t idesc1) { return (idesc1 & ~(0x1fu << 24)) | ((uint32_t)(256 >> 4) << 24); }

template <bool SUB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
coarse_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bh, GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    PairSmem &sm = *reinterpret_cast<PairSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const uint32_t p_tiles = (g.row_end - g.row0 + 2 * BM - 1) / (2 * BM);
    const uint32_t items = p_tiles * g.n_qtiles;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < PSTAGES; i++) {
            mbar_init(&sm.full[i], 1);
            mbar_init(&sm.empty[i], 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(&sm.tfull[i], 1);
            mbar_init(&sm.tempty[i], 2 * EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_bh) : "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    for (uint32_t i = threadIdx.x; i < g.n_qtiles * BN; i += blockDim.x) {
        sm.athr[i] = i < g.nq ? g.athr[i] : __int_as_float(0x7f800000);
        sm.e1[i] = (i < g.nq && g.e1) ? g.e1[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); // both CTAs' barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own rows, own half of the queries; bytes complete on the leader's barrier =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t item = cluster_id; item < items; item += n_clusters) {
                const uint32_t pt = item / g.n_qtiles, nt = item % g.n_qtiles;
                const int row = (int)(g.row0 + pt * 2 * BM + rank * BM), qrow = (int)(nt * BN + rank * (BN / 2));
                for (uint32_t kb = 0; kb < g.k_blocks; kb++) {
                    mbar_wait(&sm.empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&sm.full[stage], 2 * (A_BYTES + BH_BYTES));
                    const uint32_t full0 = mapa_u32(smem_u32(&sm.full[stage]), 0);
                    tma_load_2d_pair(sm.a[stage], &map_a, full0, (int)(kb * BK), row);
                    tma_load_2d_pair(sm.b[stage], &map_bh, full0, (int)(kb * BK), qrow);
                    if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA =====
        if (leader && lane == 0) {
            const uint32_t idesc2 = idesc_pair(g.idesc);
            uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
            for (uint32_t item = cluster_id; item < items; item += n_clusters) {
                mbar_wait(&sm.tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem + as * BN;
                for (uint32_t kb = 0; kb < g.k_blocks; kb++) {
                    mbar_wait(&sm.full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_desc(smem_u32(sm.a[stage]));
                    const uint64_t bdesc = make_desc(smem_u32(sm.b[stage]));
#pragma unroll
                    for (int k = 0; k < BK / UK; k++)
                        tc_mma_f16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc2, (kb | (uint32_t)k) != 0);
                    tc_commit_pair(&sm.empty[stage]); // frees the slot in both CTAs
                    if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(&sm.tfull[as]); // accumulators complete in both CTAs
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (both CTAs): this CTA's 128 rows x 256 queries =====
        const int ew = warp - 4;
        const uint32_t quad = (uint32_t)(warp & 3);
        const uint32_t half = (uint32_t)(ew >> 2);
        uint32_t as = 0, aphase = 0;
        for (uint32_t item = cluster_id; item < items; item += n_clusters) {
            const uint32_t pt = item / g.n_qtiles, nt = item % g.n_qtiles;
            const uint32_t row = g.row0 + pt * 2 * BM + rank * BM + quad * 32 + (uint32_t)lane;
            const bool row_ok = row < g.row_end;
            const float nr = (row_ok && g.row_l2) ? g.row_l2[row] : 0.f;
            float ra = 0.f;
            // L2: what is filtered (and stored) is a'' = a.q - ||a||^2 / 2 + c_l2 ||a||^2 — the row's own share of the L2
            // rounding term rides on the per-row constant, so it costs nothing per accumulator (rounded so a'' errs upwards)
            if constexpr (SUB) ra = row_ok ? __fmaf_rd(__fmul_ru(g.c_l2, nr), -nr, g.row_sub[row]) : 0.f;
            mbar_wait(&sm.tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (uint32_t c = 0; c < 4; c++) {
                const uint32_t col = half * 128 + c * 32;
                uint32_t r[32];
                tc_ld32(tmem + ((quad * 32) << 16) + as * BN + col, r);
                tc_wait_ld(r);
                float thr[32];
                lds_f32x32(smem_u32(&sm.athr[nt * BN + col]), thr);
                if (g.dump) {
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const uint32_t q = nt * BN + col + j;
                            if (q < g.nq) g.dump[(size_t)(row - g.row0) * g.dump_ld + q] = __uint_as_float(r[j]);
                        }
                    }
                } else {
                    uint32_t hit = 0;
                    if constexpr (SUB) {
#pragma unroll
                        for (int j = 0; j < 32; j++) r[j] = __float_as_uint(__fsub_rn(__uint_as_float(r[j]), ra));
                    }
                    float e1v[32];
                    lds_f32x32(smem_u32(&sm.e1[nt * BN + col]), e1v);
#pragma unroll
                    for (int j = 0; j < 32; j++) hit |= (__fmaf_rn(e1v[j], nr, __uint_as_float(r[j])) >= thr[j] ? 1u : 0u) << j;
                    warp_append_hits(row_ok ? hit : 0u, nt * BN + col, row, r, g.cnt, g.cand, lane, g.cand_cap);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&sm.tempty[as]), 0));
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); // nobody signals the leader's barriers or reads TMEM any more
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

// ------------------------------------------------------------------------------------------------
// mirrors and query preparation
// A lane-strided fp32 sum of d squares carries at most (d / 32 + 6) roundings of 2^-24 each (relative, all terms >= 0);
// the norms that feed the error bound are scaled up by twice that so they are upper bounds.
__host__ __device__ __forceinline__ float norm_slack(size_t dim) { return 1.0f + (float)(dim / 32 + 8) * 1.2e-7f; }

__global__ void shadow_rows_kernel(const float *__restrict__ rows, size_t row_stride_f, size_t dim, size_t first, size_t n,
                                   __nv_bfloat16 *__restrict__ shadow, size_t shadow_stride, float *__restrict__ row_l2,
                                   unsigned *__restrict__ max_l2_bits, float *__restrict__ row_hsq) {
    // one warp per row: bf16 (RNE) copy + ||row||_2 rounded up
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (size_t i = warp; i < n; i += nwarps) {
        const float *src = rows + (first + i) * row_stride_f;
        __nv_bfloat16 *dst = shadow + (first + i) * shadow_stride;
        float ss = 0.f, se = 0.f;
        for (size_t e = lane; e < shadow_stride; e += 32) {
            const float v = e < dim ? src[e] : 0.f;
            const __nv_bfloat16 b = __float2bfloat16_rn(v);
            const float dv = v - __bfloat162float(b); // exact: both share the exponent range of v
            ss = fmaf(v, v, ss);
            se = fmaf(dv, dv, se);
            dst[e] = b;
        }
        for (int w = 16; w >= 1; w >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, w);
            se += __shfl_xor_sync(0xffffffffu, se, w);
        }
        if (lane == 0) {
            const float up = norm_slack(dim);
            const float nrm = sqrtf(ss) * up;
            row_l2[first + i] = nrm;
            if (row_hsq) row_hsq[first + i] = 0.5f * ss;
            atomicMax(max_l2_bits, __float_as_uint(nrm));
            // max over rows of ||a - a^|| / ||a|| (what rounding to bf16 cost, relative to the row's norm), rounded up
            if (ss > 0.f) atomicMax(max_l2_bits + 1, __float_as_uint(__fdiv_ru(sqrtf(se) * up, __fdiv_rd(sqrtf(ss), up))));
        }
    }
}

__global__ void bf16_norms_kernel(const __nv_bfloat16 *__restrict__ rows, size_t row_stride_e, size_t dim, size_t first, size_t n,
                                  float *__restrict__ row_l2, unsigned *__restrict__ max_l2_bits, int is_fp16,
                                  float *__restrict__ row_hsq) {
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (size_t i = warp; i < n; i += nwarps) {
        const __nv_bfloat16 *src = rows + (first + i) * row_stride_e;
        float ss = 0.f;
        for (size_t e = lane; e < dim; e += 32) {
            const float v = is_fp16 ? __half2float(reinterpret_cast<const __half *>(src)[e]) : __bfloat162float(src[e]);
            ss = fmaf(v, v, ss);
        }
        for (int w = 16; w >= 1; w >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, w);
        if (lane == 0) {
            const float nrm = sqrtf(ss) * norm_slack(dim);
            if (row_l2) row_l2[first + i] = nrm;
            if (row_hsq) row_hsq[first + i] = 0.5f * ss;
            atomicMax(max_l2_bits, __float_as_uint(nrm));
        }
    }
}

// queries -> bf16 operand matrix [nq][qb_stride] (zero padded) + per-query error coefficients:
//   a^.q^ - a.q = (a^ - a).q^ + a.(q^ - q)  =>  |coarse - exact| <= ||q^|| ||a^ - a|| + ||a|| ||q^ - q||   (Cauchy-Schwarz)
//                                                               <= (||q^|| rho + ||q^ - q|| + c_rel ||q||) ||a|| = e1[q] ||a||
// with rho = max over rows of ||a^ - a|| / ||a|| and c_rel ||q|| ||a|| for the fp32 accumulation on both sides. The
// rounding terms are MEASURED (the mirror kernel records rho, this kernel ||q - q^||), not the worst case 2^-8 per
// operand: for real-valued data they are ~0.4 of it, and for stores that already hold 16-bit rows they vanish. The bound
// is per row (e1[q] times THAT row's norm): a few long rows do not widen the band of the short ones.
// e2[q]: L2 only, the part that does not scale with the row.
__global__ void prep_coarse_queries_kernel(const uint8_t *__restrict__ q, size_t q_stride, int is_f32, size_t dim, size_t nq,
                                           __nv_bfloat16 *__restrict__ qb, size_t qb_stride, float c_rel,
                                           const unsigned *__restrict__ max_l2_bits, float *__restrict__ eps, float c_l2,
                                           float *__restrict__ eps2) {
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (size_t i = warp; i < nq; i += nwarps) {
        const uint8_t *src = q + i * q_stride;
        float ss = 0.f, sb = 0.f, se = 0.f; // ||q||^2, ||q^||^2, ||q - q^||^2
        for (size_t e = lane; e < qb_stride; e += 32) {
            float v = 0.f, vb = 0.f;
            __nv_bfloat16 b = __float2bfloat16_rn(0.f);
            if (e < dim) {
                if (is_f32 == 1) {
                    v = reinterpret_cast<const float *>(src)[e];
                    b = __float2bfloat16_rn(v);
                    vb = __bfloat162float(b);
                } else if (is_f32 == 2) { // fp16 store: the 16-bit pattern is the operand
                    b = reinterpret_cast<const __nv_bfloat16 *>(src)[e];
                    v = vb = __half2float(reinterpret_cast<const __half *>(src)[e]);
                } else {
                    b = reinterpret_cast<const __nv_bfloat16 *>(src)[e];
                    v = vb = __bfloat162float(b);
                }
            }
            const float dv = v - vb;
            ss = fmaf(v, v, ss);
            sb = fmaf(vb, vb, sb);
            se = fmaf(dv, dv, se);
            qb[i * qb_stride + e] = b;
        }
        for (int w = 16; w >= 1; w >>= 1) {
            ss += __shfl_xor_sync(0xffffffffu, ss, w);
            sb += __shfl_xor_sync(0xffffffffu, sb, w);
            se += __shfl_xor_sync(0xffffffffu, se, w);
        }
        if (lane == 0) {
            // L2 adds the rounding of ||a||^2 / 2, of the subtraction and of the reference's own fp32 sum of squared
            // differences, all within c_l2 (|q| + R)^2
            const float up = norm_slack(dim);
            const float qn = sqrtf(ss) * up, qbn = sqrtf(sb) * up, qe = sqrtf(se) * up;
            const float R = __uint_as_float(max_l2_bits[0]), rho = __uint_as_float(max_l2_bits[1]);
            (void)R;
            float e = __fmaf_ru(qbn, rho, __fmaf_ru(c_rel, qn, qe));
            // L2: c_l2 (||q|| + ||a||)^2 = c_l2 ||q||^2 (e2, per query) + 2 c_l2 ||q|| ||a|| (joins e1) + c_l2 ||a||^2 (per row)
            if (eps2) {
                e = __fmaf_ru(__fmul_ru(2.0f * c_l2, qn), 1.0f, e);
                eps2[i] = __fmul_ru(__fmul_ru(c_l2, qn), qn);
            }
            eps[i] = e;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// per-query running top-k state and the merge of a phase's (exactly re-scored) candidates
__device__ __forceinline__ uint32_t f2key(float v) {
    uint32_t u = __float_as_uint(v);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
    if (k == 0xffffffffu) return __uint_as_float(0x7fc00000u);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr uint32_t RUN_CAP = 1024;   // survivors kept per query between phases (k <= 384)

struct MergeArgs {
    uint32_t nq, k;
    uint2 *run;              // [nq][RUN_CAP] (row id, coarse acc bits) of rows that can still make the top-k
    uint32_t *run_cnt;       // [nq]
    uint32_t *cnt;           // [nq] candidates appended this phase (reset to 0 here)
    const uint2 *cand;       // [nq][cand_cap]
    uint32_t run_cap, cand_cap;
    const float *e1;         // [nq] |coarse acc - exact sum| <= e1[q] * ||row|| (+ e2[q])
    const float *e2;         // [nq] or NULL (L2: the part of the bound that does not scale with the row)
    const float *row_l2;     // [rows] ||row||, rounded up
    float c_l2;              // L2: per-row share c_l2 ||row||^2 of the bound (already added to the stored value), else 0
    float *athr;             // [nq] admission bound for the next phase: admit when acc + e1 ||row|| >= athr
    uint32_t *overflow;      // [nq] set when a buffer overflowed: the query is redone exactly
    unsigned long long *total_cand;
    uint32_t m;              // sharded, phased calls: also report the bound of the m-th best row, m = ceil(k / shards); else 0
    float *bounds;           // [2 nq] out when m: bounds[q] = athr[q], bounds[nq + q] = -(bound of the m-th best), see below
};

// One block per query over (survivors U this phase's candidates). Row r's exact sum lies in [lo_r, hi_r] =
// acc_r -+ (e1 ||r|| + e2). With L_K the k-th largest lo over the rows seen so far, the k-th largest exact sum is >= L_K, so
// every row of the exact top-k — including every row tied with the k-th exact score — has hi_r >= L_K:
// that is both the survivor cut and the next phase's admission test. Only L_K is needed, not an order: an 8-bit radix
// select over (key - min key) finds it in at most four histogram passes over shared memory (the first version sorted all
// <= 4096 pairs bitonically, 78 block-wide stages: 130 us per phase at 1024 queries); the survivors are then compacted
// in any order (the re-rank orders the final list by exact score and id).
constexpr int MERGE_THREADS = 256;
__global__ void __launch_bounds__(MERGE_THREADS) merge_phase_kernel(MergeArgs a) {
    extern __shared__ __align__(16) uint8_t merge_smem[];
    const uint32_t RUN_CAP = a.run_cap, CAND_CAP = a.cand_cap;
    uint32_t *s_key = reinterpret_cast<uint32_t *>(merge_smem);        // [run + cand] ~key(lo): ascending key = descending lo
    float *s_hi = reinterpret_cast<float *>(s_key + RUN_CAP + CAND_CAP); // [run + cand] acc + e1 ||row|| (rounded up)
    uint32_t *s_rid = reinterpret_cast<uint32_t *>(s_hi + RUN_CAP + CAND_CAP); // run[] is rewritten in place: stage ids, accs
    uint32_t *s_racc = s_rid + RUN_CAP;
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_red[2 * (MERGE_THREADS / 32)];
    __shared__ uint32_t s_digit, s_below, s_keep;
    const uint32_t q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t cnt = a.cnt[q];
    const uint32_t rcnt = a.run_cnt[q];
    if (threadIdx.x == 0 && cnt) atomicAdd(a.total_cand, (unsigned long long)min(cnt, CAND_CAP));
    if (cnt > CAND_CAP) {
        if (threadIdx.x == 0) a.overflow[q] = 1;
        cnt = CAND_CAP;
    }
    const uint32_t total = rcnt + cnt;
    const uint2 *run_q = a.run + (size_t)q * RUN_CAP;
    const uint2 *cand_q = a.cand + (size_t)q * CAND_CAP;
    const float e1 = a.e1[q], e2 = a.e2 ? a.e2[q] : 0.f;
    uint32_t kmin = 0xffffffffu, kmax = 0;
    for (uint32_t i = threadIdx.x; i < total; i += MERGE_THREADS) {
        uint2 e;
        if (i < rcnt) {
            e = run_q[i];
            s_rid[i] = e.x;
            s_racc[i] = e.y;
        } else {
            e = cand_q[i - rcnt];
        }
        // stored value: acc (IP) or a'' = a' + c_l2 ||row||^2 (L2). exact sum in [v - w - 2 w2, v + w] (+- e2)
        const float nr = a.row_l2[e.x];
        const float acc = __uint_as_float(e.y), w = __fmul_ru(e1, nr), w2 = __fmul_ru(__fmul_ru(2.0f * a.c_l2, nr), nr);
        s_hi[i] = __fadd_ru(acc, w);
        const uint32_t key = ~f2key(__fsub_rd(__fsub_rd(acc, w), w2));
        s_key[i] = key;
        kmin = min(kmin, key);
        kmax = max(kmax, key);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) {
        s_red[2 * warp] = kmin;
        s_red[2 * warp + 1] = kmax;
    }
    if (threadIdx.x == 0) s_keep = 0;
    __syncthreads();
    float bound = -__int_as_float(0x7f800000), bound_m = bound;
#pragma unroll
    for (int w = 0; w < MERGE_THREADS / 32; w++) {
        kmin = min(kmin, s_red[2 * w]);
        kmax = max(kmax, s_red[2 * w + 1]);
    }
    const uint32_t range = kmax - kmin;
    const int passes = range ? (32 - __clz(range) + 7) / 8 : 0;
    // bound (in the admission test's units) of the (rank+1)-th largest lower end
    auto select = [&](uint32_t rank) -> float {
        uint32_t prefix = 0;      // digits of (rank-th smallest key - kmin) decided so far
        uint32_t want = rank;     // 0-based rank among the keys that share those digits
        for (int shift = 8 * (passes - 1); shift >= 0; shift -= 8) {
            s_hist[threadIdx.x] = 0; // MERGE_THREADS == 256 bins
            __syncthreads();
            const bool first = shift == 8 * (passes - 1); // no digits decided yet (and shift + 8 may be 32)
            for (uint32_t i = threadIdx.x; i < total; i += MERGE_THREADS) {
                const uint32_t d = s_key[i] - kmin;
                if (first || (d >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(d >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (warp == 0) {
                uint32_t h[8], sum = 0;
#pragma unroll
                for (int b = 0; b < 8; b++) {
                    h[b] = s_hist[8 * lane + b];
                    sum += h[b];
                }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                uint32_t below = incl - sum;
                if (want >= below && want < incl) {
#pragma unroll
                    for (int b = 0; b < 8; b++) {
                        if (want >= below && want < below + h[b]) {
                            s_digit = (uint32_t)(8 * lane + b);
                            s_below = below;
                        }
                        below += h[b];
                    }
                }
            }
            __syncthreads();
            prefix |= s_digit << shift;
            want -= s_below;
        }
        const float lk = key2f(~(kmin + prefix));              // (rank+1)-th largest lower end (before e2)
        float b = __fsub_rd(lk, __fmul_ru(2.0f, e2));          // hi_r + e2 >= L_K - e2
        return __fsub_rd(b, fabsf(b) * 1e-6f);                 // the epilogue's FFMA rounds to nearest
    };
    if (total >= a.k) bound = select(a.k - 1);
    // Sharded, phased call: with m = ceil(k / shards), every shard holds m rows whose exact score is at
    // least its own m-th lower end, so shards * m >= k rows reach the smallest of those: the k-th best score overall does
    // too. The caller reduces both halves of `bounds` with MAX over the shards; max(bounds[q], -bounds[nq + q]) is then a
    // valid admission bound on every shard, and far tighter than any shard's own k-th.
    if (a.m && a.m < a.k && total >= a.m) bound_m = select(a.m - 1);
    else if (a.m >= a.k) bound_m = bound;
    bound = fmaxf(bound, a.athr[q]);                           // what earlier phases (or other shards) established
    // survivors, in any order
    uint2 *run_out = a.run + (size_t)q * RUN_CAP;
    for (uint32_t i = threadIdx.x; i < total; i += MERGE_THREADS) {
        if (s_hi[i] >= bound) {
            const uint32_t slot = atomicAdd(&s_keep, 1u);
            if (slot < RUN_CAP) run_out[slot] = i < rcnt ? make_uint2(s_rid[i], s_racc[i]) : cand_q[i - rcnt];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t keep = s_keep;
        if (keep > RUN_CAP) {
            a.overflow[q] = 1;
            keep = RUN_CAP;
        }
        a.run_cnt[q] = keep;
        a.cnt[q] = 0;
        a.athr[q] = bound;
        if (a.m) {
            a.bounds[q] = bound;
            a.bounds[a.nq + q] = -bound_m;
        }
    }
}

// reduced bounds of all shards -> this shard's admission bounds for its next phase
__global__ void apply_bounds_kernel(float *__restrict__ athr, const float *__restrict__ bounds, uint32_t nq) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) athr[q] = fmaxf(athr[q], fmaxf(bounds[q], -bounds[nq + q]));
}

__global__ void unpack_ids_kernel(const uint2 *__restrict__ run, size_t total, uint32_t *__restrict__ ids) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) ids[i] = run[i].x;
}

template <typename T> __global__ void fill_t(T *p, T v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// host side
static EncodeTiledFn encode_fn() { return tc_encode_fn(); }

static int make_map(CUtensorMap *map, const void *base, size_t rows, size_t dim_elems, size_t stride_bytes, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available");
        return VSGPU_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)dim_elems, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
        return VSGPU_ERR_CUDA;
    }
    return VSGPU_OK;
}

struct TensorState {
    size_t mirrored = 0;         // rows [0, mirrored) have a valid mirror / norm
    size_t cap = 0;              // capacity the mirrors were sized for
    unsigned *max_l2_bits = nullptr;
    float *row_hsq = nullptr;     // L2 stores: ||row||^2 / 2 per row
    int sms = 0;
    bool attr_set = false, pair_attr_set = false;
    // phased top-k of a sharded caller (vsgpu_topk_device_begin / _next / _finish): what the remaining phases and the
    // re-rank need
    struct Split {
        bool armed = false;
        const uint8_t *qp = nullptr;
        size_t nq = 0, q_stride = 0, k = 0, n_ev = 0, next = 0;
        uint32_t run_cap = 0;
        uint2 *run = nullptr;
        uint32_t *rcnt = nullptr, *rid = nullptr;
        float *rsc = nullptr, *e1 = nullptr, *athr = nullptr;
        uint32_t *out_ids = nullptr;
        void *out_scores = nullptr;
        uint64_t *out_labels = nullptr;
        const float *q_norms = nullptr;
        std::vector<std::pair<uint32_t, uint32_t>> phases;
        CUtensorMap map_a, map_b, map_bh;
        GemmArgs g;
        MergeArgs m;
        size_t merge_smem = 0;
    } split;
};

// survivors whose upper score estimate does not reach the (cross-shard) bound cannot be in the global result: drop them
__global__ void __launch_bounds__(256) prune_run_kernel(uint2 *__restrict__ run, uint32_t *__restrict__ run_cnt, uint32_t run_cap,
                                                        const float *__restrict__ e1, const float *__restrict__ row_l2,
                                                        const float *__restrict__ bounds) {
    extern __shared__ uint2 s_keep[];
    __shared__ uint32_t s_n;
    const uint32_t q = blockIdx.x;
    const uint32_t cnt = run_cnt[q];
    const float b = fmaxf(bounds[q], -bounds[gridDim.x + q]), e = e1[q];
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    uint2 *mine = run + (size_t)q * run_cap;
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const uint2 v = mine[i];
        const float hi = __fadd_ru(__uint_as_float(v.y), __fmul_ru(e, row_l2[v.x]));
        if (hi >= b) s_keep[atomicAdd(&s_n, 1u)] = v;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < s_n; i += blockDim.x) mine[i] = s_keep[i];
    if (threadIdx.x == 0) run_cnt[q] = s_n;
}

static TensorState *state(vsgpu_store *s) {
    if (!s->tmap_cache) s->tmap_cache = new TensorState();
    return (TensorState *)s->tmap_cache;
}

void tensor_release(vsgpu_store *s) {
    if (s->type == VSGPU_INT8 || s->type == VSGPU_UINT8) {
        tensor_i8_release(s);
        return;
    }
    auto *t = (TensorState *)s->tmap_cache;
    if (t) {
        if (t->max_l2_bits) cudaFree(t->max_l2_bits);
        if (t->row_hsq) cudaFree(t->row_hsq);
        delete t;
        s->tmap_cache = nullptr;
    }
    if (s->shadow) cudaFree(s->shadow);
    if (s->row_l2) cudaFree(s->row_l2);
    s->shadow = nullptr;
    s->row_l2 = nullptr;
    s->shadow_stride = 0;
}

// One row changed: keep the mirror / norms in step instead of rebuilding them for the whole store (r1 rebuilt 15 GB of mirror
// after any delete). `src` != SIZE_MAX: row `src` was copied over row `id` (delete-by-swap), else row `id` was rewritten.
// The store-wide maxima (max ||row||, max relative rounding) only ever grow, so they stay valid upper bounds.
int tensor_row_changed(vsgpu_store *s, size_t id, size_t src) {
    if (s->type == VSGPU_INT8 || s->type == VSGPU_UINT8) return tensor_i8_row_changed(s, id, src);
    auto *t = (TensorState *)s->tmap_cache;
    if (!t) return VSGPU_OK;
    if (id < t->mirrored) {
        const bool f32 = s->type == VSGPU_FLOAT32;
        if (src != (size_t)-1 && src < t->mirrored) {
            if (f32 && s->shadow)
                VS_CUDA(cudaMemcpyAsync(s->shadow + id * s->shadow_stride, s->shadow + src * s->shadow_stride, s->shadow_stride * 2,
                                        cudaMemcpyDeviceToDevice, s->stream));
            if (s->row_l2) VS_CUDA(cudaMemcpyAsync(s->row_l2 + id, s->row_l2 + src, 4, cudaMemcpyDeviceToDevice, s->stream));
            if (t->row_hsq) VS_CUDA(cudaMemcpyAsync(t->row_hsq + id, t->row_hsq + src, 4, cudaMemcpyDeviceToDevice, s->stream));
        } else if (f32 && s->shadow) {
            shadow_rows_kernel<<<1, 32, 0, s->stream>>>((const float *)s->rows, s->row_stride / 4, s->dim, id, 1, (__nv_bfloat16 *)s->shadow,
                                                       s->shadow_stride, s->row_l2, t->max_l2_bits, t->row_hsq);
            VS_CUDA(cudaGetLastError());
        } else if (!f32) {
            bf16_norms_kernel<<<1, 32, 0, s->stream>>>((const __nv_bfloat16 *)s->rows, s->row_stride / 2, s->dim, id, 1, s->row_l2,
                                                      t->max_l2_bits, s->type == VSGPU_FLOAT16 ? 1 : 0, t->row_hsq);
            VS_CUDA(cudaGetLastError());
        }
    }
    t->mirrored = std::min(t->mirrored, s->count);
    return VSGPU_OK;
}

static bool g_tensor_disabled = false;

bool tensor_path_supported(const vsgpu_store *s, size_t nq, size_t k) {
    if (g_tensor_disabled) return false;
    if (s->type == VSGPU_INT8 || s->type == VSGPU_UINT8) return tensor_i8_supported(s, nq, k);
    if (s->type != VSGPU_FLOAT32 && s->type != VSGPU_BFLOAT16 && s->type != VSGPU_FLOAT16) return false;
    if (s->type == VSGPU_FLOAT16 && s->plan.kind != CK_LANES) return false; // dim >= 16: the fp32-accumulating tier
    if (s->plan.kind == CK_SEQ) return false;
    if (nq < 8 || k > K_MAX || k == 0) return false; // from 8 queries on one pass over the 16-bit rows beats the SIMT scan
    if (s->dim < 64 || s->dim > 8192) return false;
    if (s->count < 32768 || s->count < 32 * k) return false;
    if (!encode_fn()) return false;
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, s->device);
    return v == 10;
}

// bring the bf16 mirror (fp32 stores) and the row norms up to date with the store
int tensor_sync_mirrors(vsgpu_store *s) {
    TensorState *t = state(s);
    if (!t->sms) {
        cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, s->device);
        if (t->sms <= 0) t->sms = 148;
    }
    if (!t->max_l2_bits) {
        VS_CUDA(cudaMalloc(&t->max_l2_bits, 2 * sizeof(unsigned))); // [0] max ||row||, [1] max ||row - mirror row||
        VS_CUDA(cudaMemsetAsync(t->max_l2_bits, 0, 2 * sizeof(unsigned), s->stream));
    }
    if (s->metric == VSGPU_L2 && (!t->row_hsq || t->cap != s->capacity)) {
        if (t->row_hsq) cudaFree(t->row_hsq);
        t->row_hsq = nullptr;
        VS_CUDA(cudaMalloc(&t->row_hsq, s->capacity * sizeof(float)));
        t->cap = s->capacity;
        t->mirrored = 0;
    }
    if (s->type == VSGPU_FLOAT32) {
        if (!s->shadow) {
            s->shadow_stride = (s->dim + 7) / 8 * 8;
            VS_CUDA(cudaMalloc(&s->shadow, s->capacity * s->shadow_stride * 2));
            VS_CUDA(cudaMalloc(&s->row_l2, s->capacity * sizeof(float)));
            t->cap = s->capacity;
            t->mirrored = 0;
        }
        if (t->mirrored < s->count) {
            const size_t n = s->count - t->mirrored;
            const unsigned blocks = (unsigned)std::min<size_t>((n + 7) / 8, (size_t)t->sms * 16);
            shadow_rows_kernel<<<blocks, 256, 0, s->stream>>>((const float *)s->rows, s->row_stride / 4, s->dim, t->mirrored, n,
                                                             (__nv_bfloat16 *)s->shadow, s->shadow_stride, s->row_l2,
                                                             t->max_l2_bits, t->row_hsq);
            VS_CUDA(cudaGetLastError());
            s->stats.kernel_launches++;
            t->mirrored = s->count;
        }
    } else {
        if (!s->row_l2) {
            VS_CUDA(cudaMalloc(&s->row_l2, s->capacity * sizeof(float)));
            t->mirrored = 0;
        }
        if (t->mirrored < s->count) {
            const size_t n = s->count - t->mirrored;
            const unsigned blocks = (unsigned)std::min<size_t>((n + 7) / 8, (size_t)t->sms * 16);
            bf16_norms_kernel<<<blocks, 256, 0, s->stream>>>((const __nv_bfloat16 *)s->rows, s->row_stride / 2, s->dim, t->mirrored,
                                                            n, s->row_l2, t->max_l2_bits, s->type == VSGPU_FLOAT16 ? 1 : 0, t->row_hsq);
            VS_CUDA(cudaGetLastError());
            s->stats.kernel_launches++;
            t->mirrored = s->count;
        }
    }
    return VSGPU_OK;
}

static size_t al256(size_t v) { return (v + 255) / 256 * 256; }

static bool gemm_pair_enabled() {
    static const bool on = [] {
        const char *e = getenv("VSGPU_GEMM_PAIR");
        return e && e[0] == '1';
    }();
    return on;
}

// mbh: the query matrix as [128 x 64] boxes (the CTA-pair kernel stages half a query tile per CTA); may be null when
// VSGPU_GEMM_PAIR is off
static int launch_gemm(vsgpu_store *s, TensorState *t, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap *mbh,
                       GemmArgs &g) {
    if (gemm_pair_enabled() && mbh) {
        const size_t psmem = sizeof(PairSmem) + 1024;
        if (!t->pair_attr_set) {
            VS_CUDA(cudaFuncSetAttribute(coarse_gemm_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            VS_CUDA(cudaFuncSetAttribute(coarse_gemm_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            t->pair_attr_set = true;
        }
        const uint32_t p_tiles = (g.row_end - g.row0 + 2 * BM - 1) / (2 * BM);
        const uint32_t items = p_tiles * g.n_qtiles;
        const unsigned clusters = (unsigned)std::min<uint32_t>(items, (uint32_t)(t->sms / 2));
        if (g.row_sub) coarse_gemm_pair_kernel<true><<<2 * clusters, GEMM_THREADS, psmem, s->stream>>>(ma, *mbh, g);
        else coarse_gemm_pair_kernel<false><<<2 * clusters, GEMM_THREADS, psmem, s->stream>>>(ma, *mbh, g);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
        return VSGPU_OK;
    }
    const size_t smem = sizeof(GemmSmem) + 1024;
    if (!t->attr_set) {
        VS_CUDA(cudaFuncSetAttribute(coarse_gemm_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VS_CUDA(cudaFuncSetAttribute(coarse_gemm_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        t->attr_set = true;
    }
    const uint32_t m_tiles = (g.row_end - g.row0 + BM - 1) / BM;
    const uint32_t items = m_tiles * g.n_qtiles;
    const unsigned grid = (unsigned)std::min<uint32_t>(items, (uint32_t)t->sms);
    if (g.row_sub) coarse_gemm_filter_kernel<true><<<grid, GEMM_THREADS, smem, s->stream>>>(ma, mb, g);
    else coarse_gemm_filter_kernel<false><<<grid, GEMM_THREADS, smem, s->stream>>>(ma, mb, g);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

static int tensor_rerank(vsgpu_store *s, TensorState *t, const uint8_t *qp, size_t nq, size_t q_stride, size_t k, uint32_t run_cap,
                         uint2 *run, uint32_t *rcnt, uint32_t *rid, float *rsc, uint32_t *out_ids, void *out_scores,
                         uint64_t *out_labels) {
    (void)t;
    // exact scores of the survivors (bit-identical to the CPU reference), then the final order
    unpack_ids_kernel<<<256, 256, 0, s->stream>>>(run, nq * (size_t)run_cap, rid);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    VS_TRY(launch_exact_gather(s, qp, nq, q_stride, nullptr, rid, run_cap, rcnt, run_cap, rsc, run_cap));
    VS_TRY(launch_sort_candidates(s, nq, k, rid, rsc, run_cap, rcnt, k, out_ids, out_scores, out_labels));
    return VSGPU_OK;
}

// one phase of the filtered GEMM and the merge of what it admitted
static int run_phase(vsgpu_store *s, TensorState *t, const CUtensorMap &map_a, const CUtensorMap &map_b, const CUtensorMap *map_bh,
                     GemmArgs &g, const MergeArgs &m, size_t merge_smem, std::pair<uint32_t, uint32_t> rows, size_t *n_ev) {
    if (rows.first >= rows.second) return VSGPU_OK; // a shard with fewer rows than the agreed schedule covers
    g.row0 = rows.first;
    g.row_end = rows.second;
    cudaEvent_t e0 = scan_event(s, 2 * *n_ev), e1 = scan_event(s, 2 * *n_ev + 1);
    if (!e0 || !e1) return VSGPU_ERR_CUDA;
    VS_CUDA(cudaEventRecord(e0, s->stream));
    VS_TRY(launch_gemm(s, t, map_a, map_b, map_bh, g));
    VS_CUDA(cudaEventRecord(e1, s->stream));
    ++*n_ev;
    merge_phase_kernel<<<m.nq, MERGE_THREADS, merge_smem, s->stream>>>(m);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

// ph != NULL (phased call of a sharded index, single chunk of queries only): run the first phase only, hand out this
// shard's bounds and leave the rest to tensor_topk_next / tensor_topk_finish
int tensor_topk(vsgpu_store *s, const void *q_dev, size_t nq_all, size_t q_stride, const float *q_norms, size_t k,
                uint32_t *out_ids, void *out_scores, uint64_t *out_labels, const PhasedCall *ph) {
    if (s->type == VSGPU_INT8 || s->type == VSGPU_UINT8)
        return tensor_i8_topk(s, q_dev, nq_all, q_stride, q_norms, k, out_ids, out_scores, out_labels);
    TensorState *t = state(s);
    t->split.armed = false;
    const bool phased = ph != nullptr && ph->bounds != nullptr && ph->rounds > 0 && nq_all <= MAX_NQ;
    VS_TRY(tensor_sync_mirrors(s));
    const size_t n = s->count;
    const bool f32 = s->type == VSGPU_FLOAT32;
    const void *a_base = f32 ? (const void *)s->shadow : (const void *)s->rows;
    const size_t a_stride = f32 ? s->shadow_stride * 2 : s->row_stride;
    const size_t qb_stride = (s->dim + 7) / 8 * 8;
    // accumulation part of the coarse score's error bound, relative to ||q|| * max||row||: fp32 sums of d
    // terms on both sides (tensor core and reference), with slack; the operand-rounding part is measured, see
    // prep_coarse_queries_kernel
    const float c_rel = (float)((f32 ? 4.0 : s->type == VSGPU_FLOAT16 ? 8.0 : 6.0) * (double)s->dim / 8388608.0);
    const float c_l2 = s->metric == VSGPU_L2 ? (float)((double)s->dim / 4194304.0) : 0.f; // d * 2^-22
    CUtensorMap map_a;
    VS_TRY(make_map(&map_a, a_base, n, s->dim, a_stride, BM));

    const uint32_t run_cap = k <= K_SMALL_MAX ? RUN_CAP : RUN_CAP_BIG, cand_cap = k <= K_SMALL_MAX ? CAND_CAP : CAND_CAP_BIG;
    const std::vector<std::pair<uint32_t, uint32_t>> phases =
        phased ? make_phases(n, k, cand_cap, BM, ph->world, ph->rounds) : make_phases(n, k, cand_cap, BM);
    const size_t merge_smem = (size_t)(run_cap + cand_cap) * 8 + (size_t)run_cap * 8;
    if (merge_smem > 48 * 1024)
        VS_CUDA(cudaFuncSetAttribute(merge_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem));
    const size_t chunks = (nq_all + MAX_NQ - 1) / MAX_NQ;
    uint32_t *ovf_all = nullptr;
    unsigned long long *tot_all = nullptr;
    VS_TRY(pending_begin(s, nq_all, chunks, &ovf_all, &tot_all));
    size_t n_ev = 0;

    for (size_t q0 = 0; q0 < nq_all; q0 += MAX_NQ) {
        const size_t nq = std::min<size_t>(MAX_NQ, nq_all - q0);
        const uint8_t *qp = (const uint8_t *)q_dev + q0 * q_stride;
        const size_t nq_pad = (nq + BN - 1) / BN * BN;
        // scratch layout in s->cand
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
        const size_t o_qb = take(nq_pad * qb_stride * 2), o_eps = take(nq * 4), o_eps2 = take(nq * 4), o_athr = take(nq * 4), o_cnt = take(nq * 4),
                     o_rcnt = take(nq * 4), o_run = take(nq * (size_t)run_cap * 8), o_cand = take(nq * (size_t)cand_cap * 8),
                     o_rid = take(nq * (size_t)run_cap * 4), o_rsc = take(nq * (size_t)run_cap * 4);
        VS_TRY(ensure_scratch(s, s->cand, off));
        uint8_t *base = (uint8_t *)s->cand.ptr;
        auto *qb = (__nv_bfloat16 *)(base + o_qb);
        float *eps = (float *)(base + o_eps), *athr = (float *)(base + o_athr);
        float *eps2 = s->metric == VSGPU_L2 ? (float *)(base + o_eps2) : nullptr;
        uint32_t *cnt = (uint32_t *)(base + o_cnt), *ovf = ovf_all + q0, *rcnt = (uint32_t *)(base + o_rcnt);
        uint2 *run = (uint2 *)(base + o_run), *cand = (uint2 *)(base + o_cand);
        uint32_t *rid = (uint32_t *)(base + o_rid);
        float *rsc = (float *)(base + o_rsc);
        unsigned long long *tot = tot_all + q0 / MAX_NQ;

        VS_CUDA(cudaMemsetAsync(base + o_qb, 0, nq_pad * qb_stride * 2, s->stream));
        // cnt, rcnt are adjacent 256-byte-aligned blocks: clear them in one go
        VS_CUDA(cudaMemsetAsync(cnt, 0, (size_t)((uint8_t *)run - (uint8_t *)cnt), s->stream));
        fill_t<float><<<64, 256, 0, s->stream>>>(athr, -INFINITY, nq);
        prep_coarse_queries_kernel<<<(unsigned)std::min<size_t>((nq + 7) / 8, 1024), 256, 0, s->stream>>>(
            qp, q_stride, f32 ? 1 : (s->type == VSGPU_FLOAT16 ? 2 : 0), s->dim, nq, qb, qb_stride, c_rel, t->max_l2_bits, eps, c_l2, eps2);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches += 2;
        CUtensorMap map_b, map_bh;
        VS_TRY(make_map(&map_b, qb, nq_pad, s->dim, qb_stride * 2, BN));
        if (gemm_pair_enabled()) VS_TRY(make_map(&map_bh, qb, nq_pad, s->dim, qb_stride * 2, BN / 2));

        GemmArgs g{};
        g.nq = (uint32_t)nq;
        g.n_qtiles = (uint32_t)(nq_pad / BN);
        g.k_blocks = (uint32_t)((s->dim + BK - 1) / BK);
        g.athr = athr;
        g.e1 = eps;
        g.row_l2 = s->row_l2;
        g.c_l2 = c_l2;
        g.cnt = cnt;
        g.cand = cand;
        g.cand_cap = cand_cap;
        g.dump = nullptr;
        g.idesc = s->type == VSGPU_FLOAT16 ? IDESC_FP16 : IDESC_BF16;
        g.row_sub = s->metric == VSGPU_L2 ? t->row_hsq : nullptr;
        MergeArgs m{};
        m.nq = (uint32_t)nq;
        m.k = (uint32_t)k;
        m.run = run;
        m.run_cnt = rcnt;
        m.cnt = cnt;
        m.cand = cand;
        m.run_cap = run_cap;
        m.cand_cap = cand_cap;
        m.e1 = eps;
        m.e2 = eps2;
        m.row_l2 = s->row_l2;
        m.c_l2 = c_l2;
        m.athr = athr;
        m.overflow = ovf;
        m.total_cand = tot;
        if (phased) {
            m.m = (uint32_t)((k + ph->world - 1) / ph->world);
            m.bounds = ph->bounds;
            VS_TRY(run_phase(s, t, map_a, map_b, gemm_pair_enabled() ? &map_bh : nullptr, g, m, merge_smem, phases[0], &n_ev));
            auto &sp = t->split;
            sp.armed = true;
            sp.qp = qp;
            sp.nq = nq;
            sp.q_stride = q_stride;
            sp.k = k;
            sp.n_ev = n_ev;
            sp.next = 1;
            sp.run_cap = run_cap;
            sp.run = run;
            sp.rcnt = rcnt;
            sp.rid = rid;
            sp.rsc = rsc;
            sp.e1 = eps;
            sp.athr = athr;
            sp.out_ids = out_ids;
            sp.out_scores = out_scores;
            sp.out_labels = out_labels;
            sp.q_norms = q_norms;
            sp.phases = phases;
            sp.map_a = map_a;
            sp.map_b = map_b;
            sp.map_bh = map_bh;
            sp.g = g;
            sp.m = m;
            sp.merge_smem = merge_smem;
            return VSGPU_OK;
        }
        for (size_t p = 0; p < phases.size(); p++)
            VS_TRY(run_phase(s, t, map_a, map_b, gemm_pair_enabled() ? &map_bh : nullptr, g, m, merge_smem, phases[p], &n_ev));
        VS_TRY(tensor_rerank(s, t, qp, nq, q_stride, k, run_cap, run, rcnt, rid, rsc, out_ids ? out_ids + q0 * k : nullptr,
                             out_scores ? (float *)out_scores + q0 * k : nullptr, out_labels ? out_labels + q0 * k : nullptr));
    }
    // overflowed queries are redone on the exact path by whoever synchronises next (no host wait here)
    VS_TRY(pending_arm(s, q_dev, nq_all, q_stride, q_norms, k, out_ids, out_scores, out_labels, n_ev, chunks));
    return VSGPU_OK;
}

// a phased call that was begun and never finished must not leak its phases into the next one
void tensor_topk_reset(vsgpu_store *s) {
    if (s->type == VSGPU_INT8 || s->type == VSGPU_UINT8) return;
    if (auto *t = (TensorState *)s->tmap_cache) t->split.armed = false;
}

// phased call, after the caller reduced `bounds` over the shards: tighten this shard's admission bounds, run its next phase
// and write its new bounds. A no-op once the phases are done (or when _begin did all the work on another path).
int tensor_topk_next(vsgpu_store *s, float *bounds) {
    auto *t = (TensorState *)s->tmap_cache;
    if (s->type == VSGPU_INT8 || s->type == VSGPU_UINT8 || !t || !t->split.armed) return VSGPU_OK;
    auto &sp = t->split;
    if (sp.next >= sp.phases.size()) return VSGPU_OK;
    const auto rows = sp.phases[sp.next++];
    if (rows.first >= rows.second) return VSGPU_OK;
    apply_bounds_kernel<<<(unsigned)((sp.nq + 255) / 256), 256, 0, s->stream>>>(sp.athr, bounds, (uint32_t)sp.nq);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    sp.m.bounds = bounds;
    return run_phase(s, t, sp.map_a, sp.map_b, gemm_pair_enabled() ? &sp.map_bh : nullptr, sp.g, sp.m, sp.merge_smem, rows, &sp.n_ev);
}

// last step of a phased call: any phases left (a caller that exchanged fewer times than agreed), prune with the reduced
// bounds (may be NULL), re-rank, write the lists
int tensor_topk_finish(vsgpu_store *s, const float *bounds) {
    auto *t = (TensorState *)s->tmap_cache;
    if (s->type == VSGPU_INT8 || s->type == VSGPU_UINT8 || !t || !t->split.armed) return VSGPU_OK;
    auto &sp = t->split;
    sp.armed = false;
    if (bounds && sp.next < sp.phases.size()) {
        apply_bounds_kernel<<<(unsigned)((sp.nq + 255) / 256), 256, 0, s->stream>>>(sp.athr, bounds, (uint32_t)sp.nq);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
    }
    const bool ran_more = sp.next < sp.phases.size();
    sp.m.m = 0; // nobody reads the bounds of these phases
    for (; sp.next < sp.phases.size(); sp.next++)
        VS_TRY(run_phase(s, t, sp.map_a, sp.map_b, gemm_pair_enabled() ? &sp.map_bh : nullptr, sp.g, sp.m, sp.merge_smem,
                         sp.phases[sp.next], &sp.n_ev));
    if (bounds && !ran_more) {
        prune_run_kernel<<<(unsigned)sp.nq, 256, (size_t)sp.run_cap * sizeof(uint2), s->stream>>>(sp.run, sp.rcnt, sp.run_cap, sp.e1,
                                                                                                s->row_l2, bounds);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
    }
    VS_TRY(tensor_rerank(s, t, sp.qp, sp.nq, sp.q_stride, sp.k, sp.run_cap, sp.run, sp.rcnt, sp.rid, sp.rsc, sp.out_ids, sp.out_scores,
                         sp.out_labels));
    VS_TRY(pending_arm(s, sp.qp, sp.nq, sp.q_stride, sp.q_norms, sp.k, sp.out_ids, sp.out_scores, sp.out_labels, sp.n_ev, 1));
    return VSGPU_OK;
}

size_t tensor_topk_rounds(size_t rows, size_t k, unsigned world) {
    return topk_rounds(rows, k, k <= K_SMALL_MAX ? CAND_CAP : CAND_CAP_BIG, BM, world);
}

} // namespace vsgpu

// Host-logic test hook (no device needed): the phase schedule of the filtered GEMM for n rows and top-k. Writes up to `cap`
// phase end rows to `edges`, returns the number of phases.
extern "C" size_t vsgpu_debug_phases_sharded(size_t n, size_t k, unsigned world, size_t rounds, uint32_t *edges, size_t cap) {
    const auto ph = vsgpu::make_phases(n, k, k <= vsgpu::K_SMALL_MAX ? vsgpu::CAND_CAP : vsgpu::CAND_CAP_BIG, vsgpu::BM, world, rounds);
    for (size_t i = 0; i < ph.size() && i < cap; i++) edges[i] = ph[i].second;
    return ph.size();
}
extern "C" size_t vsgpu_debug_phases(size_t n, size_t k, uint32_t *edges, size_t cap) {
    const auto ph = vsgpu::make_phases(n, k, vsgpu::CAND_CAP, vsgpu::BM);
    for (size_t i = 0; i < ph.size() && i < cap; i++) edges[i] = ph[i].second;
    return ph.size();
}

// Debug / test hook: raw coarse accumulators of rows [row0, row0 + nrows) against nq queries
// (HOST pointers; out is [nrows][nq] fp32). Exercises exactly the production TMA/MMA/TMEM pipeline.
extern "C" int vsgpu_debug_coarse(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t row0, size_t nrows,
                                  float *out) {
    using namespace vsgpu;
    if (!s || !queries || !out || nq == 0 || nrows == 0 || row0 % BM != 0 || row0 + nrows > s->count || nq > MAX_NQ) {
        set_error("vsgpu_debug_coarse: bad arguments");
        return VSGPU_ERR_ARG;
    }
    if (s->type != VSGPU_FLOAT32 && s->type != VSGPU_BFLOAT16 && s->type != VSGPU_FLOAT16) {
        set_error("vsgpu_debug_coarse: fp32 / bf16 / fp16 stores only");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    TensorState *t = state(s);
    VS_TRY(tensor_sync_mirrors(s));
    const bool f32 = s->type == VSGPU_FLOAT32;
    const size_t qb_stride = (s->dim + 7) / 8 * 8;
    const size_t nq_pad = (nq + BN - 1) / BN * BN;
    uint8_t *d_q = nullptr;
    __nv_bfloat16 *qb = nullptr;
    float *eps = nullptr, *athr = nullptr, *dump = nullptr;
    VS_CUDA(cudaMalloc(&d_q, nq * s->row_bytes));
    VS_CUDA(cudaMalloc(&qb, nq_pad * qb_stride * 2));
    VS_CUDA(cudaMalloc(&eps, nq * 4));
    VS_CUDA(cudaMalloc(&athr, nq * 4));
    VS_CUDA(cudaMalloc(&dump, nrows * nq * 4));
    VS_CUDA(cudaMemcpy2D(d_q, s->row_bytes, queries, qstride, s->row_bytes, nq, cudaMemcpyHostToDevice));
    // a copy from pageable memory may return before its DMA has landed, and the store's stream does not wait for the
    // default stream
    VS_CUDA(cudaDeviceSynchronize());
    VS_CUDA(cudaMemsetAsync(qb, 0, nq_pad * qb_stride * 2, s->stream));
    VS_CUDA(cudaMemsetAsync(dump, 0, nrows * nq * 4, s->stream));
    prep_coarse_queries_kernel<<<64, 256, 0, s->stream>>>(d_q, s->row_bytes, f32 ? 1 : (s->type == VSGPU_FLOAT16 ? 2 : 0), s->dim, nq, qb, qb_stride, 0.f,
                                                         t->max_l2_bits, eps, 0.f, nullptr);
    CUtensorMap map_a, map_b;
    VS_TRY(make_map(&map_a, f32 ? (const void *)s->shadow : (const void *)s->rows, s->count, s->dim,
                    f32 ? s->shadow_stride * 2 : s->row_stride, BM));
    VS_TRY(make_map(&map_b, qb, nq_pad, s->dim, qb_stride * 2, BN));
    CUtensorMap map_bh;
    if (gemm_pair_enabled()) VS_TRY(make_map(&map_bh, qb, nq_pad, s->dim, qb_stride * 2, BN / 2));
    GemmArgs g{};
    g.row0 = (uint32_t)row0;
    g.row_end = (uint32_t)(row0 + nrows);
    g.nq = (uint32_t)nq;
    g.n_qtiles = (uint32_t)(nq_pad / BN);
    g.k_blocks = (uint32_t)((s->dim + BK - 1) / BK);
    g.athr = athr;
    g.cnt = nullptr;
    g.cand = nullptr;
    g.dump = dump;
    g.dump_ld = (uint32_t)nq;
    g.idesc = s->type == VSGPU_FLOAT16 ? IDESC_FP16 : IDESC_BF16;
    g.row_sub = nullptr;
    VS_TRY(launch_gemm(s, t, map_a, map_b, gemm_pair_enabled() ? &map_bh : nullptr, g));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    VS_CUDA(cudaMemcpy(out, dump, nrows * nq * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_q);
    cudaFree(qb);
    cudaFree(eps);
    cudaFree(athr);
    cudaFree(dump);
    return VSGPU_OK;
}
