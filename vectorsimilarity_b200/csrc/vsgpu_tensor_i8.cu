// Tensor-core path for int8 / uint8 stores (BASELINE configs[2]: int8 Cosine, d = 512, batch 4096).
//
// For 8-bit integers the batched scan is an EXACT integer GEMM: tcgen05.mma kind::i8 accumulates
// sum(a_i * b_i) in int32 in TMEM, which is precisely the integer the reference's VNNI kernels compute
// (spaces/IP/IP_AVX512F_BW_VL_VNNI_INT8.h:27-77, L2/L2_AVX512F_BW_VL_VNNI_INT8.h:28-65, the uint8
// twins). So there is no coarse pass and no re-rank here: the epilogue turns "score <= T_q" into a
// test on the accumulator (dot >= c_q * row_mul + row_add, loosened by a few ulps so the admitted
// set is a superset), the few admitted (row, dot) pairs are scored with the reference's exact final
// formula — float(1 - dot), float(aa + qq - 2 dot), 1.0f - float(dot) / (norm_a * norm_b) with .rn
// intrinsics (vsgpu_dist.cuh int_score) — and merged into the running top-k by (score, id).
//
// Pipeline = the bf16 kernel's (vsgpu_tensor.cu): TMA 128B-swizzled [128 x 128 B] row boxes and
// [256 x 128 B] query boxes, 4-stage mbarrier ring, one MMA-issuing thread, 128 x 256 int32
// accumulator double-buffered in TMEM, 8 epilogue warps. Rows are scanned in geometric phases; after
// each phase one block per query merges the phase's candidates and tightens T_q.
#include "vsgpu_dist.cuh"
#include "vsgpu_tc.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace vsgpu {

namespace {

constexpr int BM = 128;           // rows per tile
constexpr int BN = 256;           // queries per tile
constexpr int BKB = 128;          // bytes (= int8 elements) per k-block: one 128-byte swizzle row
constexpr int UKB = 32;           // bytes per tcgen05.mma kind::i8 (K = 32)
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BKB; // 16 KB
constexpr int B_BYTES = BN * BKB; // 32 KB
constexpr int MAX_NQ = 4096;
// 16 epilogue warps (four per TMEM lane quadrant, 64 accumulator columns each): with 8 the epilogue — TMEM load latency,
// I2F + FFMA + FSETP per accumulator, the append's atomic round trip — was the critical path at two warps per scheduler
// (13.6 ms against 9.2 ms with the epilogue compiled out, profiles/r2_i8_epilogue_ab.md)
constexpr int EPI_WARPS = 16;
constexpr int EPI_COLS = 4 * 256 / EPI_WARPS; // accumulator columns per epilogue warp
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
constexpr uint32_t CAND_CAP = 3072; // candidates per query and phase
constexpr uint32_t RUN_CAP = 1024;  // k <= RUN_CAP

struct I8Smem {
    uint8_t a[STAGES][A_BYTES];
    uint8_t b[STAGES][B_BYTES];
    float cq[MAX_NQ];
    float cqmin[MAX_NQ / 32]; // the smallest admission constant of each 32-query chunk: the quick reject below
    uint64_t full[STAGES], empty[STAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
};

struct I8Args {
    uint32_t row0, row_end, nq, n_qtiles, k_blocks, idesc;
    uint32_t mma_only; // measurement: the epilogue only drains TMEM (no test, no append) -> the pipeline's own ceiling
    const float *cq;      // [nq] admit when float(dot) >= cq * row_mul + row_add (loosened)
    const float *row_mul; // per row, or NULL = 1 (cosine: the stored norm)
    const float *row_add; // per row, or NULL = 0 (L2: sum a^2 / 2)
    uint32_t *cnt;        // [nq]
    uint2 *cand;          // [nq][CAND_CAP] (row id, dot)
    int *dump;            // debug: [rows][dump_ld] raw accumulators
    uint32_t dump_ld;
};

__global__ void __launch_bounds__(GEMM_THREADS, 1)
i8_gemm_filter_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, I8Args g) {
    extern __shared__ uint8_t smem_raw[];
    I8Smem &sm = *reinterpret_cast<I8Smem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
    for (uint32_t i = threadIdx.x; i < g.n_qtiles * BN; i += blockDim.x)
        sm.cq[i] = (i < g.nq && g.cq) ? g.cq[i] : __int_as_float(0x7f800000);
    __syncthreads();
    for (uint32_t grp = (uint32_t)warp; grp < g.n_qtiles * BN / 32; grp += blockDim.x / 32) {
        float m = sm.cq[grp * 32 + (uint32_t)lane];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) sm.cqmin[grp] = m;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) { // ===== TMA producer =====
            uint32_t stage = 0, phase = 0;
            for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
                const uint32_t mt = item / g.n_qtiles, nt = item % g.n_qtiles;
                const int row = (int)(g.row0 + mt * BM), qrow = (int)(nt * BN);
                for (uint32_t kb = 0; kb < g.k_blocks; kb++) {
                    mbar_wait(&sm.empty[stage], phase ^ 1);
                    mbar_expect_tx(&sm.full[stage], A_BYTES + B_BYTES);
                    tma_load_2d(sm.a[stage], &map_a, &sm.full[stage], (int)(kb * BKB), row);
                    tma_load_2d(sm.b[stage], &map_b, &sm.full[stage], (int)(kb * BKB), qrow);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) { // ===== MMA issuer =====
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
                    for (int k = 0; k < BKB / UKB; k++) // 32 bytes (2 x 16 B) along K inside the swizzled row
                        tc_mma_i8(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), g.idesc, (kb | (uint32_t)k) != 0);
                    tc_commit(&sm.empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit(&sm.tfull[as]);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM int32 -> bound test -> (row, dot) candidates =====
        const int ew = warp - 4;
        const uint32_t quad = (uint32_t)(warp & 3);
        const uint32_t half = (uint32_t)(ew >> 2); // which EPI_COLS-wide slice of the 256 columns
        uint32_t as = 0, aphase = 0;
        // the per-row terms of the NEXT tile are fetched while this one is tested: a global load in front of every tile's
        // first compare was ~a fifth of the tile time (long-scoreboard stalls, profiles/r2_i8_gemm_ncu.md)
        auto row_terms = [&](uint32_t item, float &rm_o, float &ra_o) {
            const uint32_t r = g.row0 + (item / g.n_qtiles) * BM + quad * 32 + (uint32_t)lane;
            const bool ok = item < items && r < g.row_end;
            rm_o = (ok && g.row_mul) ? __ldg(g.row_mul + r) : 1.0f;
            ra_o = (ok && g.row_add) ? __ldg(g.row_add + r) : 0.0f;
        };
        float rm_n, ra_n;
        row_terms(blockIdx.x, rm_n, ra_n);
        DeferredHits dh;
        dh.init();
        for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
            const uint32_t mt = item / g.n_qtiles, nt = item % g.n_qtiles;
            const uint32_t row = g.row0 + mt * BM + quad * 32 + (uint32_t)lane;
            const bool row_ok = row < g.row_end;
            // admit when float(dot) >= c_q * rm + ra2: c_q arrives already lowered by the merge kernel, ra2 is the
            // row's additive term lowered by a few ulps and two dot units (the admitted set must be a superset)
            const float rm = rm_n, ra = ra_n;
            row_terms(item + gridDim.x, rm_n, ra_n);
            const float ra2 = ra - fabsf(ra) * 3.8146973e-6f - 2.0f;
            mbar_wait(&sm.tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (uint32_t c = 0; c < EPI_COLS / 32; c++) {
                const uint32_t col = half * EPI_COLS + c * 32;
                uint32_t r[32];
                tc_ld32(tmem + ((quad * 32) << 16) + as * BN + col, r);
                tc_wait_ld(r);
                if (!g.dump && g.mma_only == 0) {
                    // Quick reject: a hit needs float(dot_j) >= fma(c_j, rm, ra2) for some j, so the row's largest dot must
                    // reach the bound built from the chunk's smallest c_j (rm >= 0, fma and int -> float are monotone).
                    // 11 integer max3 + one convert / fma / compare instead of four instructions per accumulator; in the
                    // long late phases practically every chunk stops here — the compare loop was the busiest pipe of
                    // this kernel (ALU 62 %, tensor 54 %: profiles/r2_i8_gemm_final_ncu.md).
                    int m = (int)r[0];
#pragma unroll
                    for (int j = 1; j < 32; j++) m = max(m, (int)r[j]);
                    const float tmin = fmaf(sm.cqmin[(nt * BN + col) >> 5], rm, ra2);
                    if (!__any_sync(0xffffffffu, row_ok && __int2float_rn(m) >= tmin)) continue;
                }
                float cq[32];
                lds_f32x32(smem_u32(&sm.cq[nt * BN + col]), cq);
                if (g.dump) {
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const uint32_t q = nt * BN + col + j;
                            if (q < g.nq) g.dump[(size_t)(row - g.row0) * g.dump_ld + q] = (int)r[j];
                        }
                    }
                } else if (g.mma_only == 1) {
                    uint32_t x = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) x ^= r[j];
                    if (x == 0xdeadbeefu && cq[0] == 12345.f) g.cnt[0] = x; // keeps the loads alive, never true in practice
                } else {
                    // 3 instructions per accumulator (I2F, FFMA, FSETP), all 32 first, then the hits of the whole warp;
                    // padded query columns carry c_q = +inf
                    uint32_t hit = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) hit |= (__int2float_rn((int)r[j]) >= fmaf(cq[j], rm, ra2) ? 1u : 0u) << j;
                    if (g.mma_only == 2) { // measurement: the test without the append
                        if (hit == 0xdeadbeefu && cq[0] == 12345.f) g.cnt[0] = hit;
                    } else {
                        warp_append_hits_deferred(row_ok ? hit : 0u, nt * BN + col, row, r, g.cnt, g.cand, lane, CAND_CAP, dh);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&sm.tempty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
        dh.flush(g.cand, CAND_CAP);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
}

// ---- CTA-pair variant: tcgen05.mma.cta_group::2, a 256-row x 256-query tile per cluster of two ----------------------
// The single-CTA kernel moves (128 + 256) x 128 operand bytes per k-block for 128 x 256 x 128 MACs and runs at the L2 -> SM
// bandwidth cap, far below the int8 MMA rate. Here each CTA stages its own 128 rows and HALF of the query
// tile: (128 + 128) x 128 bytes for the same MACs per CTA — a third less operand traffic per MAC. The leader (cluster rank
// 0) issues the MMAs; both CTAs' operands complete on the leader's `full` barriers; commits are multicast to both CTAs.
constexpr int PSTAGES = 6;
constexpr int BH_BYTES = (BN / 2) * BKB; // 16 KB: this CTA's half of the query tile
struct I8PairSmem {
    uint8_t a[PSTAGES][A_BYTES];
    uint8_t b[PSTAGES][BH_BYTES];
    float cq[MAX_NQ];
    uint64_t full[PSTAGES], empty[PSTAGES], tfull[2], tempty[2];
    uint32_t tmem_base;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
i8_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bh, I8Args g) {
    extern __shared__ uint8_t smem_raw[];
    I8PairSmem &sm = *reinterpret_cast<I8PairSmem *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
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
    for (uint32_t i = threadIdx.x; i < g.n_qtiles * BN; i += blockDim.x)
        sm.cq[i] = (i < g.nq && g.cq) ? g.cq[i] : __int_as_float(0x7f800000);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all(); // both CTAs' barriers exist before anything is signalled across the pair
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) { // ===== TMA producer (both CTAs): own rows, own half of the queries =====
            uint32_t stage = 0, phase = 0;
            for (uint32_t item = cluster_id; item < items; item += n_clusters) {
                const uint32_t pt = item / g.n_qtiles, nt = item % g.n_qtiles;
                const int row = (int)(g.row0 + pt * 2 * BM + rank * BM), qrow = (int)(nt * BN + rank * (BN / 2));
                for (uint32_t kb = 0; kb < g.k_blocks; kb++) {
                    mbar_wait(&sm.empty[stage], phase ^ 1);
                    if (leader) mbar_expect_tx(&sm.full[stage], 2 * (A_BYTES + BH_BYTES));
                    const uint32_t full0 = mapa_u32(smem_u32(&sm.full[stage]), 0);
                    tma_load_2d_pair(sm.a[stage], &map_a, full0, (int)(kb * BKB), row);
                    tma_load_2d_pair(sm.b[stage], &map_bh, full0, (int)(kb * BKB), qrow);
                    if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) { // ===== MMA issuer: one thread of the leader CTA =====
            const uint32_t idesc2 = (g.idesc & ~(0x1fu << 24)) | ((uint32_t)(256 >> 4) << 24);
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
                    for (int k = 0; k < BKB / UKB; k++)
                        tc_mma_i8_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc2, (kb | (uint32_t)k) != 0);
                    tc_commit_pair(&sm.empty[stage]);
                    if (++stage == PSTAGES) { stage = 0; phase ^= 1; }
                }
                tc_commit_pair(&sm.tfull[as]);
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
            const float rm = (row_ok && g.row_mul) ? g.row_mul[row] : 1.0f;
            const float ra = (row_ok && g.row_add) ? g.row_add[row] : 0.0f;
            const float ra2 = ra - fabsf(ra) * 3.8146973e-6f - 2.0f;
            mbar_wait(&sm.tfull[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (uint32_t c = 0; c < EPI_COLS / 32; c++) {
                const uint32_t col = half * EPI_COLS + c * 32;
                uint32_t r[32];
                tc_ld32(tmem + ((quad * 32) << 16) + as * BN + col, r);
                tc_wait_ld(r);
                float cq[32];
                lds_f32x32(smem_u32(&sm.cq[nt * BN + col]), cq);
                if (g.dump) {
                    if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const uint32_t q = nt * BN + col + j;
                            if (q < g.nq) g.dump[(size_t)(row - g.row0) * g.dump_ld + q] = (int)r[j];
                        }
                    }
                } else if (g.mma_only) {
                    uint32_t x = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) x ^= r[j];
                    if (x == 0xdeadbeefu && cq[0] == 12345.f) g.cnt[0] = x;
                } else {
                    uint32_t hit = 0;
#pragma unroll
                    for (int j = 0; j < 32; j++) hit |= (__int2float_rn((int)r[j]) >= fmaf(cq[j], rm, ra2) ? 1u : 0u) << j;
                    warp_append_hits(row_ok ? hit : 0u, nt * BN + col, row, r, g.cnt, g.cand, lane, CAND_CAP);
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

// ---- per-row / per-query integer side data --------------------------------------------------------
// row_sq[i] = sum a^2 (exact, u32: 255^2 * dim < 2^32 for dim < 66 051), row_add[i] = row_sq / 2 as float
__global__ void i8_row_sq_kernel(const uint8_t *__restrict__ rows, size_t row_stride, int chunks, int is_unsigned, size_t first,
                                 size_t n, uint32_t *__restrict__ row_sq, float *__restrict__ row_add) {
    const size_t gid = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const size_t ng = ((size_t)gridDim.x * blockDim.x) >> 3;
    const int c = threadIdx.x & 7;
    for (size_t i = gid; i < (n + ng - 1) / ng * ng; i += ng) {
        const bool ok = i < n;
        unsigned long long t = 0;
        if (ok) {
            const uint4 *rp = reinterpret_cast<const uint4 *>(rows + (first + i) * row_stride);
            for (int ch = c; ch < chunks; ch += 8) {
                const uint4 v = __ldg(rp + ch);
                if (is_unsigned) t += (unsigned)dot4<true>(v.x, v.x, 0) + (unsigned)dot4<true>(v.y, v.y, 0) + (unsigned)dot4<true>(v.z, v.z, 0) + (unsigned)dot4<true>(v.w, v.w, 0);
                else t += dot4<false>(v.x, v.x, 0) + dot4<false>(v.y, v.y, 0) + dot4<false>(v.z, v.z, 0) + dot4<false>(v.w, v.w, 0);
            }
        }
        for (int w = 4; w >= 1; w >>= 1) t += __shfl_xor_sync(0xffffffffu, t, w);
        if (ok && c == 0) {
            row_sq[first + i] = (uint32_t)t;
            row_add[first + i] = 0.5f * (float)t;
        }
    }
}
__global__ void i8_query_sq_kernel(const uint8_t *__restrict__ q, size_t q_stride, int chunks, int is_unsigned, size_t nq,
                                   long long *__restrict__ qq) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 *qp = reinterpret_cast<const uint4 *>(q + i * q_stride);
        long long t = 0;
        for (int ch = 0; ch < chunks; ch++) {
            const uint4 v = qp[ch];
            if (is_unsigned) t += (long long)(unsigned)dot4<true>(v.x, v.x, 0) + (unsigned)dot4<true>(v.y, v.y, 0) + (long long)(unsigned)dot4<true>(v.z, v.z, 0) + (unsigned)dot4<true>(v.w, v.w, 0);
            else t += (long long)dot4<false>(v.x, v.x, 0) + dot4<false>(v.y, v.y, 0) + (long long)dot4<false>(v.z, v.z, 0) + dot4<false>(v.w, v.w, 0);
        }
        qq[i] = t;
    }
}

__device__ __forceinline__ uint32_t score_key(float v) {
    uint32_t u = __float_as_uint(v);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu; // NaN last
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct I8MergeArgs {
    uint32_t nq, k;
    int metric;
    uint2 *run;         // [nq][RUN_CAP] (row id, exact score bits), ascending (score, id)
    uint32_t *run_cnt;  // [nq]
    uint32_t *cnt;      // [nq] candidates of this phase (reset here)
    const uint2 *cand;  // [nq][CAND_CAP] (row id, dot)
    const uint32_t *row_sq;
    const float *row_norm;
    const long long *qq;
    const float *q_norm;
    float *cq;          // [nq] next phase's admission constant
    uint32_t *overflow; // [nq]
    unsigned long long *total_cand;
};

// One block per query: exact scores of this phase's candidates, merge with the running top-k by
// (score, id), tighten the admission constant.
// 256 threads: seven blocks per SM (32 KB of sort space each) — with 1024 threads two fit, and a batch of 4096 queries
// took 14 waves whose block-wide barriers dominated (0.1 ms per phase, 0.5 ms for the unfiltered first phase).
constexpr int I8_MERGE_THREADS = 256;
constexpr uint32_t I8_SEL_CAP = 1024; // pairs the selection may keep (k <= 1024 and its ties); more: full sort
__global__ void __launch_bounds__(I8_MERGE_THREADS) i8_merge_kernel(I8MergeArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const uint32_t q = blockIdx.x;
    uint32_t cnt = a.cnt[q];
    const uint32_t rcnt = a.run_cnt[q];
    if (threadIdx.x == 0 && cnt) atomicAdd(a.total_cand, (unsigned long long)min(cnt, CAND_CAP));
    if (cnt > CAND_CAP) {
        if (threadIdx.x == 0) a.overflow[q] = 1;
        cnt = CAND_CAP;
    }
    const uint32_t total = rcnt + cnt;
    uint32_t P = 32;
    while (P < total) P <<= 1;
    uint32_t *sk = reinterpret_cast<uint32_t *>(smem_raw);
    uint32_t *si = sk + P;
    const long long qq = a.qq[q];
    const float qn = a.q_norm ? a.q_norm[q] : 0.f;
    for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
        uint32_t key = 0xffffffffu, id = 0xffffffffu;
        if (i < rcnt) {
            const uint2 e = a.run[(size_t)q * RUN_CAP + i];
            id = e.x;
            key = score_key(__uint_as_float(e.y));
        } else if (i < total) {
            const uint2 e = a.cand[(size_t)q * CAND_CAP + (i - rcnt)];
            id = e.x;
            const long long dot = (long long)(int)e.y;
            const long long aa = a.metric == VSGPU_L2 ? (long long)a.row_sq[id] : 0;
            const float rn = a.row_norm ? a.row_norm[id] : 0.f;
            key = score_key(int_score(a.metric, dot, aa, qq, rn, qn));
        }
        if (i < total && key == 0xffffffffu) key = 0xfffffffeu; // real entries ahead of the padding
        sk[i] = key;
        si[i] = id;
    }
    __syncthreads();
    // Only the k best are kept, so a full sort of up to 4096 pairs is wasted work (the unfiltered first phase sorted 2048
    // pairs per query: 0.43 ms for 4096 queries). Above 64 pairs: radix-select the k-th smallest key (four 8-bit passes
    // over shared memory), move the pairs up to and including it — all of its ties, the order among them is by id — to
    // the front of a second buffer and sort those few. Falls back to the full sort when the ties do not fit.
    uint32_t *ck = si + P, *ci = ck + I8_SEL_CAP;
    __shared__ uint32_t s_hist[256];
    __shared__ uint32_t s_digit, s_below, s_n;
    bool selected = false;
    if (total > 64 && total > a.k) {
        uint32_t prefix = 0, want = a.k - 1;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) s_hist[i] = 0;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
                const uint32_t key = sk[i];
                if (shift == 24 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                const int lane = threadIdx.x;
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
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            if (sk[i] <= prefix) {
                const uint32_t slot = atomicAdd(&s_n, 1u);
                if (slot < I8_SEL_CAP) {
                    ck[slot] = sk[i];
                    ci[slot] = si[i];
                }
            }
        }
        __syncthreads();
        const uint32_t c = s_n;
        if (c <= I8_SEL_CAP) {
            uint32_t P2 = 32;
            while (P2 < c) P2 <<= 1;
            for (uint32_t i = c + threadIdx.x; i < P2; i += blockDim.x) {
                ck[i] = 0xffffffffu;
                ci[i] = 0xffffffffu;
            }
            __syncthreads();
            P = P2;
            sk = ck;
            si = ci;
            selected = true;
        }
    }
    (void)selected;
    for (uint32_t size = 2; size <= P; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const uint32_t ka = sk[lo], kb = sk[hi], ia = si[lo], ib = si[hi];
                const bool gt = ka > kb || (ka == kb && ia > ib);
                if (gt == asc) { sk[lo] = kb; sk[hi] = ka; si[lo] = ib; si[hi] = ia; }
            }
            __syncthreads();
        }
    }
    const uint32_t keep = min(total, a.k);
    auto key_score = [](uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); };
    for (uint32_t i = threadIdx.x; i < keep; i += blockDim.x)
        a.run[(size_t)q * RUN_CAP + i] = make_uint2(si[i], __float_as_uint(sk[i] >= 0xfffffffeu ? __int_as_float(0x7fc00000) : key_score(sk[i])));
    if (threadIdx.x == 0) {
        a.run_cnt[q] = keep;
        a.cnt[q] = 0;
        float cq = -__int_as_float(0x7f800000); // admit everything until k rows are known
        if (total >= a.k && sk[a.k - 1] < 0xfffffffeu) {
            const float T = key_score(sk[a.k - 1]);
            // score <= T  <=>  dot >= c_q * row_mul + row_add, each loosened by a few ulps of its terms
            if (a.metric == VSGPU_COSINE) cq = qn * ((1.0f - T) - 3.8146973e-6f);          // dot >= (1 - T) |a| |q|
            else if (a.metric == VSGPU_IP) cq = (1.0f - T) - fabsf(T) * 9.5367432e-7f - 1.0f; // dot >= 1 - T
            else cq = 0.5f * ((float)qq - T) - (fabsf((float)qq) + fabsf(T)) * 4.7683716e-7f - 1.0f; // dot >= aa/2 + (qq - T)/2
        }
        // lowered once more by 2^-18 of its magnitude: covers the rounding of c_q * row_mul + row_add and of
        // float(dot) in the epilogue (row_mul > 0)
        a.cq[q] = cq - fabsf(cq) * 3.8146973e-6f;
    }
}

__global__ void i8_emit_kernel(const uint2 *__restrict__ run, const uint32_t *__restrict__ run_cnt, const uint64_t *__restrict__ labels,
                               size_t nq, size_t k, uint32_t *__restrict__ out_ids, float *__restrict__ out_scores,
                               uint64_t *__restrict__ out_labels) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq * k; i += (size_t)gridDim.x * blockDim.x) {
        const size_t q = i / k, j = i % k;
        const bool ok = j < run_cnt[q];
        const uint2 e = ok ? run[q * RUN_CAP + j] : make_uint2(0xffffffffu, 0x7fc00000u);
        if (out_ids) out_ids[i] = e.x;
        if (out_scores) out_scores[i] = __uint_as_float(e.y);
        if (out_labels) out_labels[i] = ok ? labels[e.x] : ~0ull;
    }
}

__global__ void i8_fill_kernel(float *p, float v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

struct I8State {
    size_t synced = 0; // rows [0, synced) have row_sq / row_add
    size_t cap = 0;
    uint32_t *row_sq = nullptr;
    float *row_add = nullptr;
    int sms = 0;
    bool attr_set = false, pair_attr_set = false, merge_attr_set = false;
};

int make_map_u8(CUtensorMap *map, const void *base, size_t rows, size_t dim_bytes, size_t stride_bytes, int box_rows) {
    EncodeTiledFn fn = tc_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available");
        return VSGPU_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)dim_bytes, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)stride_bytes};
    cuuint32_t box[2] = {(cuuint32_t)BKB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (u8) failed with code " + std::to_string((int)r));
        return VSGPU_ERR_CUDA;
    }
    return VSGPU_OK;
}

size_t al256(size_t v) { return (v + 255) / 256 * 256; }

I8State *i8_state(vsgpu_store *s) {
    if (!s->tmap_cache) s->tmap_cache = new I8State();
    return (I8State *)s->tmap_cache;
}

} // namespace

void tensor_i8_release(vsgpu_store *s) {
    auto *t = (I8State *)s->tmap_cache;
    if (!t) return;
    if (t->row_sq) cudaFree(t->row_sq);
    if (t->row_add) cudaFree(t->row_add);
    delete t;
    s->tmap_cache = nullptr;
}

int tensor_i8_row_changed(vsgpu_store *s, size_t id, size_t src) {
    auto *t = (I8State *)s->tmap_cache;
    if (!t) return VSGPU_OK;
    if (t->row_sq && id < t->synced) {
        if (src != (size_t)-1 && src < t->synced) {
            VS_CUDA(cudaMemcpyAsync(t->row_sq + id, t->row_sq + src, 4, cudaMemcpyDeviceToDevice, s->stream));
            VS_CUDA(cudaMemcpyAsync(t->row_add + id, t->row_add + src, 4, cudaMemcpyDeviceToDevice, s->stream));
        } else {
            i8_row_sq_kernel<<<1, 32, 0, s->stream>>>(s->rows, s->row_stride, (int)(s->row_stride / 16), s->type == VSGPU_UINT8, id, 1,
                                                     t->row_sq, t->row_add);
            VS_CUDA(cudaGetLastError());
        }
    }
    t->synced = std::min(t->synced, s->count);
    return VSGPU_OK;
}

bool tensor_i8_supported(const vsgpu_store *s, size_t nq, size_t k) {
    if (s->type != VSGPU_INT8 && s->type != VSGPU_UINT8) return false;
    if (nq < 8 || k == 0 || k > RUN_CAP) return false;
    if (s->dim < 32 || s->dim > 33024) return false; // int32 accumulator: the reference's own cap (spaces.h:57-66)
    if (s->count < 32768 || s->count < 16 * k) return false;
    if (!tc_encode_fn()) return false;
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, s->device);
    return v == 10;
}

static int i8_sync_side(vsgpu_store *s, I8State *t) {
    if (!t->sms) {
        cudaDeviceGetAttribute(&t->sms, cudaDevAttrMultiProcessorCount, s->device);
        if (t->sms <= 0) t->sms = 148;
    }
    if (s->metric != VSGPU_L2) return VSGPU_OK;
    if (t->cap < s->capacity) {
        uint32_t *nsq = nullptr;
        float *nadd = nullptr;
        VS_CUDA(cudaMalloc(&nsq, s->capacity * 4));
        VS_CUDA(cudaMalloc(&nadd, s->capacity * 4));
        if (t->row_sq) cudaFree(t->row_sq);
        if (t->row_add) cudaFree(t->row_add);
        t->row_sq = nsq;
        t->row_add = nadd;
        t->cap = s->capacity;
        t->synced = 0;
    }
    if (t->synced < s->count) {
        const size_t n = s->count - t->synced;
        const unsigned blocks = (unsigned)std::min<size_t>((n * 8 + 255) / 256, (size_t)t->sms * 16);
        i8_row_sq_kernel<<<blocks, 256, 0, s->stream>>>(s->rows, s->row_stride, (int)(s->row_stride / 16), s->type == VSGPU_UINT8, t->synced,
                                                       n, t->row_sq, t->row_add);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
        t->synced = s->count;
    }
    return VSGPU_OK;
}

static int env_flag(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
static bool i8_pair_enabled() {
    static const bool on = env_flag("VSGPU_I8_PAIR", 0) != 0; // cta_group::2 variant: opt-in (measured slower)
    return on;
}

// mbh: the query matrix as [128 x 128 B] boxes (the pair kernel stages half a query tile per CTA), NULL = single-CTA kernel
static int i8_launch_gemm(vsgpu_store *s, I8State *t, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap *mbh,
                          I8Args &g) {
    static const int mma_only = env_flag("VSGPU_I8_MMA_ONLY", 0);
    g.mma_only = g.dump ? 0u : (uint32_t)mma_only;
    if (mbh) {
        const size_t psmem = sizeof(I8PairSmem) + 1024;
        if (!t->pair_attr_set) {
            VS_CUDA(cudaFuncSetAttribute(i8_gemm_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            t->pair_attr_set = true;
        }
        const uint32_t p_tiles = (g.row_end - g.row0 + 2 * BM - 1) / (2 * BM);
        const unsigned clusters = (unsigned)std::min<uint32_t>(p_tiles * g.n_qtiles, (uint32_t)(t->sms / 2));
        i8_gemm_pair_kernel<<<2 * clusters, GEMM_THREADS, psmem, s->stream>>>(ma, *mbh, g);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
        return VSGPU_OK;
    }
    const size_t smem = sizeof(I8Smem) + 1024;
    if (!t->attr_set) {
        VS_CUDA(cudaFuncSetAttribute(i8_gemm_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        t->attr_set = true;
    }
    const uint32_t m_tiles = (g.row_end - g.row0 + BM - 1) / BM;
    const unsigned grid = (unsigned)std::min<uint32_t>(m_tiles * g.n_qtiles, (uint32_t)t->sms);
    i8_gemm_filter_kernel<<<grid, GEMM_THREADS, smem, s->stream>>>(ma, mb, g);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

static uint32_t i8_idesc(const vsgpu_store *s) {
    const uint32_t fmt = s->type == VSGPU_INT8 ? 1u : 0u; // S8: 0 unsigned, 1 signed
    // D = S32 (bits 4-5 = 2), A/B formats at bits 7-9 / 10-12, K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    return (2u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// q_dev: queries packed at the store's row stride (zero padded), q_norms for cosine (stage_queries_device)
int tensor_i8_topk(vsgpu_store *s, const void *q_dev, size_t nq_all, size_t q_stride, const float *q_norms, size_t k,
                   uint32_t *out_ids, void *out_scores, uint64_t *out_labels) {
    I8State *t = i8_state(s);
    VS_TRY(i8_sync_side(s, t));
    const size_t n = s->count;
    const bool uns = s->type == VSGPU_UINT8;
    CUtensorMap map_a;
    VS_TRY(make_map_u8(&map_a, s->rows, n, s->dim, s->row_stride, BM));
    const std::vector<std::pair<uint32_t, uint32_t>> phases = make_phases(n, k, CAND_CAP, BM);
    const size_t chunks = (nq_all + MAX_NQ - 1) / MAX_NQ;
    uint32_t *ovf_all = nullptr;
    unsigned long long *tot_all = nullptr;
    VS_TRY(pending_begin(s, nq_all, chunks, &ovf_all, &tot_all));
    cudaEvent_t e0 = scan_event(s, 0), e1 = scan_event(s, 1);
    if (!e0 || !e1) return VSGPU_ERR_CUDA;
    VS_CUDA(cudaEventRecord(e0, s->stream));
    for (size_t q0 = 0; q0 < nq_all; q0 += MAX_NQ) {
        const size_t nq = std::min<size_t>(MAX_NQ, nq_all - q0);
        const uint8_t *qp = (const uint8_t *)q_dev + q0 * q_stride;
        const float *qn = q_norms ? q_norms + q0 : nullptr;
        const size_t nq_pad = (nq + BN - 1) / BN * BN;
        size_t off = 0;
        auto take = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
        const size_t o_cq = take(nq * 4), o_cnt = take(nq * 4), o_rcnt = take(nq * 4), o_qq = take(nq * 8),
                     o_run = take(nq * RUN_CAP * 8), o_cand = take(nq * CAND_CAP * 8);
        VS_TRY(ensure_scratch(s, s->cand, off));
        uint8_t *base = (uint8_t *)s->cand.ptr;
        float *cq = (float *)(base + o_cq);
        uint32_t *cnt = (uint32_t *)(base + o_cnt), *ovf = ovf_all + q0, *rcnt = (uint32_t *)(base + o_rcnt);
        long long *qq = (long long *)(base + o_qq);
        unsigned long long *tot = tot_all + q0 / MAX_NQ;
        uint2 *run = (uint2 *)(base + o_run), *cand = (uint2 *)(base + o_cand);
        // cnt, rcnt, qq are adjacent: clear in one go; the first phase admits everything (cq = -inf)
        VS_CUDA(cudaMemsetAsync(cnt, 0, (size_t)((uint8_t *)run - (uint8_t *)cnt), s->stream));
        i8_fill_kernel<<<16, 256, 0, s->stream>>>(cq, -INFINITY, nq);
        i8_query_sq_kernel<<<(unsigned)std::min<size_t>((nq + 127) / 128, 1024), 128, 0, s->stream>>>(qp, q_stride, (int)(s->row_stride / 16),
                                                                                                  uns, nq, qq);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
        CUtensorMap map_b, map_bh;
        VS_TRY(make_map_u8(&map_b, qp, nq, s->dim, q_stride, BN));
        if (i8_pair_enabled()) VS_TRY(make_map_u8(&map_bh, qp, nq, s->dim, q_stride, BN / 2));
        for (size_t p = 0; p < phases.size(); p++) {
            I8Args g{};
            g.row0 = phases[p].first;
            g.row_end = phases[p].second;
            g.nq = (uint32_t)nq;
            g.n_qtiles = (uint32_t)(nq_pad / BN);
            g.k_blocks = (uint32_t)((s->dim + BKB - 1) / BKB);
            g.idesc = i8_idesc(s);
            g.cq = cq;
            g.row_mul = s->metric == VSGPU_COSINE ? s->norms : nullptr;
            g.row_add = s->metric == VSGPU_L2 ? t->row_add : nullptr;
            g.cnt = cnt;
            g.cand = cand;
            VS_TRY(i8_launch_gemm(s, t, map_a, map_b, i8_pair_enabled() ? &map_bh : nullptr, g));
            I8MergeArgs m{};
            m.nq = (uint32_t)nq;
            m.k = (uint32_t)k;
            m.metric = s->metric;
            m.run = run;
            m.run_cnt = rcnt;
            m.cnt = cnt;
            m.cand = cand;
            m.row_sq = t->row_sq;
            m.row_norm = s->metric == VSGPU_COSINE ? s->norms : nullptr;
            m.qq = qq;
            m.q_norm = s->metric == VSGPU_COSINE ? qn : nullptr;
            m.cq = cq;
            m.overflow = ovf;
            m.total_cand = tot;
            const size_t msm = 2 * 4096 * 4 + 2 * I8_SEL_CAP * 4;
            if (!t->merge_attr_set) { // a per-device attribute: one flag per store, not per process
                VS_CUDA(cudaFuncSetAttribute(i8_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm));
                t->merge_attr_set = true;
            }
            i8_merge_kernel<<<(unsigned)nq, I8_MERGE_THREADS, msm, s->stream>>>(m);
            VS_CUDA(cudaGetLastError());
            s->stats.kernel_launches++;
        }
        i8_emit_kernel<<<(unsigned)std::min<size_t>((nq * k + 255) / 256, 2048), 256, 0, s->stream>>>(
            run, rcnt, s->labels, nq, k, out_ids ? out_ids + q0 * k : nullptr, out_scores ? (float *)out_scores + q0 * k : nullptr,
            out_labels ? out_labels + q0 * k : nullptr);
        VS_CUDA(cudaGetLastError());
        s->stats.kernel_launches++;
    }
    VS_CUDA(cudaEventRecord(e1, s->stream));
    // queries whose candidate buffer overflowed (adversarial ties) are redone on the exact path by whoever synchronises next
    VS_TRY(pending_arm(s, q_dev, nq_all, q_stride, q_norms, k, out_ids, out_scores, out_labels, 1, chunks));
    return VSGPU_OK;
}

} // namespace vsgpu

// Debug / test hook: raw int32 accumulators of rows [row0, row0 + nrows) against nq queries (HOST pointers; queries
// are dim bytes each; out is [nrows][nq] int32). Exercises the production TMA / kind::i8 MMA / TMEM pipeline.
extern "C" int vsgpu_debug_i8(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t row0, size_t nrows, int *out) {
    using namespace vsgpu;
    if (!s || !queries || !out || nq == 0 || nrows == 0 || row0 % BM != 0 || row0 + nrows > s->count || nq > MAX_NQ ||
        (s->type != VSGPU_INT8 && s->type != VSGPU_UINT8)) {
        set_error("vsgpu_debug_i8: bad arguments");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    I8State *t = i8_state(s);
    VS_TRY(i8_sync_side(s, t));
    uint8_t *d_q = nullptr;
    int *dump = nullptr;
    VS_CUDA(cudaMalloc(&d_q, nq * s->row_stride));
    VS_CUDA(cudaMalloc(&dump, nrows * nq * 4));
    VS_CUDA(cudaMemset(d_q, 0, nq * s->row_stride));
    VS_CUDA(cudaMemcpy2D(d_q, s->row_stride, queries, qstride, s->row_bytes, nq, cudaMemcpyHostToDevice));
    // the clear and the copy ran on the default stream, which the store's (non-blocking) stream does not wait for — and a
    // copy from pageable memory may return before its DMA has landed
    VS_CUDA(cudaDeviceSynchronize());
    VS_CUDA(cudaMemsetAsync(dump, 0, nrows * nq * 4, s->stream));
    CUtensorMap map_a, map_b, map_bh;
    VS_TRY(make_map_u8(&map_a, s->rows, s->count, s->dim, s->row_stride, BM));
    VS_TRY(make_map_u8(&map_b, d_q, nq, s->dim, s->row_stride, BN));
    if (i8_pair_enabled()) VS_TRY(make_map_u8(&map_bh, d_q, nq, s->dim, s->row_stride, BN / 2));
    I8Args g{};
    g.row0 = (uint32_t)row0;
    g.row_end = (uint32_t)(row0 + nrows);
    g.nq = (uint32_t)nq;
    g.n_qtiles = (uint32_t)((nq + BN - 1) / BN);
    g.k_blocks = (uint32_t)((s->dim + BKB - 1) / BKB);
    g.idesc = i8_idesc(s);
    g.dump = dump;
    g.dump_ld = (uint32_t)nq;
    VS_TRY(i8_launch_gemm(s, t, map_a, map_b, i8_pair_enabled() ? &map_bh : nullptr, g));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    VS_CUDA(cudaMemcpy(out, dump, nrows * nq * 4, cudaMemcpyDeviceToHost));
    cudaFree(d_q);
    cudaFree(dump);
    return VSGPU_OK;
}
