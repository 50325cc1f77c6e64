// Streaming exact scan for the AVX512F "lanes" tier without residual (fp32 / fp16 stores, dim % 32
// == 0 — the benchmark shapes d = 128 / 768 / 1024): same bit-exact arithmetic as exact_scan_kernel
// (vsgpu_exact.cu; reference spaces/IP/IP_AVX512F_FP32.h:19-56, L2/L2_AVX512F_FP32.h:21-59,
// IP/IP_AVX512F_FP16.h:27-68), restructured so the HBM stream, the FMA pipe and shared memory all
// stay busy at once:
//
//   * rows never touch registers on their way in: a producer warp issues TMA bulk copies
//     (cp.async.bulk, one per row segment) of [64 rows x KC elements] tiles into a ring of
//     shared-memory stages guarded by mbarriers; 8 consumer warps drain them;
//   * a consumer warp owns 8 rows x up to 16 queries = 128 accumulators: lane c is chain c (the
//     x86 lane / accumulator pair), one LDS feeds 16 FMAs, and the FMAs are issued as packed
//     FFMA2 (fma.rn.f32x2: two independently IEEE-rounded FMAs per issue slot — bit-identical to
//     scalar fma.rn.f32) over pairs of queries;
//   * the 32-lane reduction of _mm512_reduce_add_ps is a transposing butterfly: each round halves
//     the number of live values per lane instead of reducing every value on every lane (same
//     pairing tree xor 16, 8, 4, 2, 1, fp add is commutative, so the bits are unchanged).
//
// Roofline: HBM. Algorithmic bytes per launch = n * dim * sizeof(T) + nq * n * 4 (scores out).
#include "vsgpu_dist.cuh"
#include <algorithm>
#include <cuda.h>

namespace vsgpu {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void bar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool bar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_addr(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!bar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap(); // a lost arrival must not hang the device
    }
}
// TMA tile copy global -> shared ([box rows x box columns] of the row-major store), completion counted in
// bytes on `bar`; rows past the end of the store are zero-filled by the engine
__device__ __forceinline__ void tma_tile(void *dst, const CUtensorMap *map, uint64_t *bar, int col, int row) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_addr(dst)),
                 "l"(map), "r"(smem_addr(bar)), "r"(col), "r"(row)
                 : "memory");
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

constexpr int TILE_ROWS = 64;    // rows per stage: 8 consumer warps x 8 rows
constexpr int ROWS_PER_WARP = 8;
constexpr int CONSUMER_WARPS = 8;
constexpr int SCAN_THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int MAX_STAGES = 8;

struct ScanTmaArgs {
    const uint8_t *rows;
    size_t row_stride;
    size_t n;
    int dim;
    int kc;      // elements per stage row segment (multiple of 32, divides dim)
    int nstages;
    const uint8_t *q; // raw query blobs (same element type as the rows)
    size_t q_stride;
    int nq;
    float *scores; // [nq][ld]
    size_t ld;
    // fused selection (TOPK kernels): per-CTA running top-k in shared memory, only k (key, id) pairs per CTA and query leave
    int k;          // <= FUSED_K_MAX
    uint2 *part;    // [nq][gridDim.x][k] (score key, row id), ascending (key, id); unused slots carry id = 0xffffffff
};

// Fused top-k: the k best (score, id) pairs a CTA has seen per query live in shared memory; a row is appended only when it
// beats the CTA's current k-th score, and the buffer is cut back to k by a block-wide bitonic sort whenever it fills
// (a few times per CTA: the number of appends is ~ k (1 + ln(rows per CTA / k))).
constexpr int FUSED_K_MAX = 128;
constexpr int FUSED_CAP = 512;      // buffer entries per query
constexpr int FUSED_CHECK_TILES = 4; // occupancy is checked every that many tiles (<= 64 appends per tile and query)

__device__ __forceinline__ uint32_t score_key32(float v) { // order-preserving, NaN last (as vsgpu_select.cu)
    const uint32_t u = __float_as_uint(v);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32) : "memory"); }

// sort buf[0 .. FUSED_CAP) ascending (entries are key << 32 | id; empty slots = ~0), consumer threads only
__device__ __forceinline__ void fused_sort(unsigned long long *buf) {
    for (int size = 2; size <= FUSED_CAP; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < FUSED_CAP / 2; t += CONSUMER_WARPS * 32) {
                const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const unsigned long long x = buf[lo], y = buf[hi];
                if ((x > y) == asc) {
                    buf[lo] = y;
                    buf[hi] = x;
                }
            }
            consumer_barrier();
        }
}

template <typename ET> __device__ __forceinline__ float elem_to_float(ET v);
template <> __device__ __forceinline__ float elem_to_float<float>(float v) { return v; }
template <> __device__ __forceinline__ float elem_to_float<__half>(__half v) { return __half2float(v); }

// QP = query pairs per pass (2, 4 or 8)
template <typename ET, int QP, bool L2, bool TOPK>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_tma_kernel(const __grid_constant__ CUtensorMap map, ScanTmaArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int QC = 2 * QP;
    constexpr int V = ROWS_PER_WARP * QC;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem);
    uint64_t *empty = full + MAX_STAGES;
    float2 *qs = reinterpret_cast<float2 *>(smem + 128);                           // [QP][dim]
    unsigned char *after_q = smem + 128 + (((size_t)QP * a.dim * sizeof(float2) + 127) / 128) * 128;
    // TOPK: [QC] buffers of FUSED_CAP (key << 32 | id) entries, then per query the append cursor and the admission key
    unsigned long long *tk_buf = reinterpret_cast<unsigned long long *>(after_q);
    uint32_t *tk_cnt = reinterpret_cast<uint32_t *>(after_q + (size_t)QC * FUSED_CAP * 8);
    uint32_t *tk_thr = tk_cnt + QC;
    unsigned char *stage0 = TOPK ? after_q + (size_t)QC * FUSED_CAP * 8 + 128 : after_q;
    const size_t seg_bytes = (size_t)a.kc * sizeof(ET);
    const size_t stage_bytes = (size_t)TILE_ROWS * seg_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = a.dim / a.kc;
    const int steps = a.kc / 32;

    if (threadIdx.x == 0) {
        for (int i = 0; i < a.nstages; i++) {
            bar_init(&full[i], 1);
            bar_init(&empty[i], CONSUMER_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // queries, pair-interleaved: qs[p][e] = (q[2p][e], q[2p+1][e]); absent queries are zero
    for (int i = threadIdx.x; i < QP * a.dim; i += blockDim.x) {
        const int p = i / a.dim, e = i % a.dim;
        float2 v = make_float2(0.f, 0.f);
        if (2 * p < a.nq) v.x = elem_to_float<ET>(reinterpret_cast<const ET *>(a.q + (size_t)(2 * p) * a.q_stride)[e]);
        if (2 * p + 1 < a.nq) v.y = elem_to_float<ET>(reinterpret_cast<const ET *>(a.q + (size_t)(2 * p + 1) * a.q_stride)[e]);
        qs[i] = v;
    }
    if constexpr (TOPK) {
        for (int i = threadIdx.x; i < QC * FUSED_CAP; i += blockDim.x) tk_buf[i] = ~0ull;
        if (threadIdx.x < QC) {
            tk_cnt[threadIdx.x] = 0;
            tk_thr[threadIdx.x] = 0xffffffffu; // admit everything until k rows are known
        }
    }
    __syncthreads();

    const size_t ntiles = (a.n + TILE_ROWS - 1) / TILE_ROWS;
    if (warp == CONSUMER_WARPS) {
        // ---- producer: one TMA tile copy per (row tile, column chunk), issued by one elected lane ----
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map) : "memory");
            uint32_t it = 0;
            for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int ch = 0; ch < nchunks; ch++, it++) {
                    const int st = it % a.nstages;
                    const uint32_t ph = (it / a.nstages) & 1;
                    bar_wait(&empty[st], ph ^ 1);
                    bar_expect_tx(&full[st], (uint32_t)stage_bytes);
                    tma_tile(stage0 + (size_t)st * stage_bytes, &map, &full[st], ch * a.kc, (int)(tile * TILE_ROWS));
                }
            }
        }
        return;
    }

    // ---- consumers ----
    uint32_t it = 0;
    // TOPK: cut query q's buffer back to its k best when it could overflow before the next check (block-uniform decision)
    auto compact = [&](bool force) {
        consumer_barrier(); // every append of the tiles so far is visible
        for (int q = 0; q < a.nq; q++) {
            const uint32_t c = tk_cnt[q];
            if (!force && c <= (uint32_t)(FUSED_CAP - 64 * FUSED_CHECK_TILES)) continue;
            unsigned long long *buf = tk_buf + (size_t)q * FUSED_CAP;
            consumer_barrier();
            fused_sort(buf);
            const uint32_t keep = min(c, (uint32_t)a.k);
            for (int i = (int)keep + threadIdx.x; i < FUSED_CAP; i += CONSUMER_WARPS * 32) buf[i] = ~0ull;
            if (threadIdx.x == 0) {
                tk_cnt[q] = keep;
                // rows are visited in ascending id order within a CTA: a later row that only ties the k-th score loses
                if (keep == (uint32_t)a.k) tk_thr[q] = (uint32_t)(buf[a.k - 1] >> 32);
            }
            consumer_barrier();
        }
    };
    int tiles_done = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        u64 acc[ROWS_PER_WARP][QP];
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; r++)
#pragma unroll
            for (int p = 0; p < QP; p++) acc[r][p] = 0ull;
        for (int ch = 0; ch < nchunks; ch++, it++) {
            const int st = it % a.nstages;
            const uint32_t ph = (it / a.nstages) & 1;
            bar_wait(&full[st], ph);
            const ET *xb = reinterpret_cast<const ET *>(stage0 + (size_t)st * stage_bytes + (size_t)(warp * ROWS_PER_WARP) * seg_bytes) + lane;
            const float2 *qb = qs + (size_t)ch * a.kc + lane;
#pragma unroll 2
            for (int s = 0; s < steps; s++) {
                u64 qv[QP];
#pragma unroll
                for (int p = 0; p < QP; p++) qv[p] = *reinterpret_cast<const u64 *>(qb + (size_t)p * a.dim + s * 32);
#pragma unroll
                for (int r = 0; r < ROWS_PER_WARP; r++) {
                    const float x = elem_to_float<ET>(xb[(size_t)r * a.kc + s * 32]);
                    const u64 xx = pack2(x, x);
#pragma unroll
                    for (int p = 0; p < QP; p++) {
                        if constexpr (L2) {
                            const u64 d = sub2(xx, qv[p]);
                            acc[r][p] = fma2(d, d, acc[r][p]);
                        } else {
                            acc[r][p] = fma2(xx, qv[p], acc[r][p]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) bar_arrive(&empty[st]);
        }
        // ---- reduce over the 32 chains: transposing butterfly (xor 16, 8, 4, 2, 1) ----
        float v[V];
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; r++)
#pragma unroll
            for (int p = 0; p < QP; p++) unpack2(acc[r][p], v[r * QC + 2 * p], v[r * QC + 2 * p + 1]);
        {
            int half = V / 2;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const bool up = (lane & m) != 0;
#pragma unroll
                for (int i = 0; i < V / 2; i++) {
                    if (i < half) {
                        const float lo = v[i], hi = v[i + half];
                        const float send = up ? lo : hi, keep = up ? hi : lo;
                        v[i] = __fadd_rn(keep, __shfl_xor_sync(0xffffffffu, send, m));
                    }
                }
                half >>= 1;
            }
        }
        // lane L now holds value indices L * (V/32) + j: row L >> 2, queries (L & 3) * (QC/4) + j
        constexpr int PER = V / 32;
        const size_t row = tile * TILE_ROWS + (size_t)warp * ROWS_PER_WARP + (lane >> 2);
        if constexpr (TOPK) {
            if (row < a.n) {
#pragma unroll
                for (int j = 0; j < PER; j++) {
                    const int q = (lane & 3) * PER + j;
                    if (q < a.nq) {
                        const uint32_t key = score_key32(L2 ? v[j] : __fsub_rn(1.0f, v[j]));
                        if (key < tk_thr[q] || tk_thr[q] == 0xffffffffu) {
                            const uint32_t slot = atomicAdd(&tk_cnt[q], 1u);
                            if (slot < (uint32_t)FUSED_CAP) tk_buf[(size_t)q * FUSED_CAP + slot] = ((unsigned long long)key << 32) | (uint32_t)row;
                        }
                    }
                }
            }
            if (++tiles_done % FUSED_CHECK_TILES == 0) compact(false);
        } else {
            if (row < a.n) {
#pragma unroll
                for (int j = 0; j < PER; j++) {
                    const int q = (lane & 3) * PER + j;
                    if (q < a.nq) a.scores[(size_t)q * a.ld + row] = L2 ? v[j] : __fsub_rn(1.0f, v[j]);
                }
            }
        }
    }
    if constexpr (TOPK) {
        compact(true);
        for (int i = threadIdx.x; i < a.nq * a.k; i += CONSUMER_WARPS * 32) {
            const int q = i / a.k, j = i % a.k;
            const unsigned long long e = tk_buf[(size_t)q * FUSED_CAP + j];
            a.part[((size_t)q * gridDim.x + blockIdx.x) * a.k + j] = make_uint2((uint32_t)(e >> 32), e == ~0ull ? 0xffffffffu : (uint32_t)e);
        }
    }
}

// Merge of the per-CTA lists (one block per query): the k-th best overall is no worse than any single CTA's k-th, so only
// entries up to T = min over CTAs of their k-th key can matter; they are gathered into shared memory FUSED_MERGE_CAP at a
// time, sorted by (key, id) and cut back to k — any number of ties at T is handled by repeating.
constexpr int FUSED_MERGE_CAP = 2048;
template <int THREADS>
__global__ void __launch_bounds__(THREADS) fused_merge_kernel(const uint2 *__restrict__ part, int parts, int k, const uint64_t *__restrict__ labels,
                                                              uint32_t out_ld, uint32_t *__restrict__ out_ids, float *__restrict__ out_scores,
                                                              uint64_t *__restrict__ out_labels) {
    __shared__ unsigned long long buf[FUSED_MERGE_CAP];
    __shared__ uint32_t s_T, s_n;
    const int q = blockIdx.x;
    const uint2 *mine = part + (size_t)q * parts * k;
    if (threadIdx.x == 0) {
        s_T = 0xffffffffu;
        s_n = 0;
    }
    for (int i = threadIdx.x; i < FUSED_MERGE_CAP; i += THREADS) buf[i] = ~0ull;
    __syncthreads();
    for (int p = threadIdx.x; p < parts; p += THREADS) {
        const uint2 last = mine[(size_t)p * k + (k - 1)];
        if (last.y != 0xffffffffu) atomicMin(&s_T, last.x); // this CTA alone holds k rows at or below its k-th key
    }
    __syncthreads();
    const uint32_t T = s_T;
    const int total = parts * k;
    auto sort_and_cut = [&]() {
        for (int size = 2; size <= FUSED_MERGE_CAP; size <<= 1)
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = threadIdx.x; t < FUSED_MERGE_CAP / 2; t += THREADS) {
                    const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                    const bool asc = (lo & size) == 0;
                    const unsigned long long x = buf[lo], y = buf[hi];
                    if ((x > y) == asc) {
                        buf[lo] = y;
                        buf[hi] = x;
                    }
                }
                __syncthreads();
            }
        for (int i = k + threadIdx.x; i < FUSED_MERGE_CAP; i += THREADS) buf[i] = ~0ull;
        if (threadIdx.x == 0) s_n = min(s_n, (uint32_t)k);
        __syncthreads();
    };
    for (int base = 0; base < total; base += THREADS) {
        const int i = base + threadIdx.x;
        bool take = false;
        uint2 e = make_uint2(0, 0);
        if (i < total) {
            e = mine[i];
            take = e.y != 0xffffffffu && e.x <= T;
        }
        // room for a whole round of THREADS entries, else sort and cut first (block-uniform)
        if (s_n + THREADS > FUSED_MERGE_CAP) sort_and_cut();
        if (take) {
            const uint32_t slot = atomicAdd(&s_n, 1u);
            buf[slot] = ((unsigned long long)e.x << 32) | e.y;
        }
        __syncthreads();
    }
    sort_and_cut();
    for (uint32_t i = threadIdx.x; i < out_ld; i += THREADS) {
        const unsigned long long e = i < (uint32_t)k ? buf[i] : ~0ull;
        const bool ok = e != ~0ull;
        const uint32_t id = ok ? (uint32_t)e : 0xffffffffu, key = (uint32_t)(e >> 32);
        if (out_ids) out_ids[(size_t)q * out_ld + i] = id;
        if (out_scores) {
            const float sc = !ok || key == 0xffffffffu ? __uint_as_float(0x7fc00000u)
                                                       : __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
            out_scores[(size_t)q * out_ld + i] = sc;
        }
        if (out_labels) out_labels[(size_t)q * out_ld + i] = ok ? labels[id] : ~0ull;
    }
}

int sm_count(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    return v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static const EncodeTiledFn fn = [] { // initialised once, also under concurrent first calls (sharded index threads)
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        EncodeTiledFn f = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            f = (EncodeTiledFn)p;
        cudaGetLastError();
        return f;
    }();
    return fn;
}

template <typename ET, bool L2, bool TOPK = false> int launch_t(vsgpu_store *s, ScanTmaArgs &a, size_t smem_bytes, unsigned grid = 0) {
    CUtensorMap map;
    {
        cuuint64_t gdim[2] = {(cuuint64_t)a.dim, (cuuint64_t)a.n};
        cuuint64_t gstride[1] = {(cuuint64_t)a.row_stride};
        cuuint32_t box[2] = {(cuuint32_t)a.kc, (cuuint32_t)TILE_ROWS};
        cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode_fn()(&map, sizeof(ET) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                                       const_cast<uint8_t *>(a.rows), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled (scan) failed with code " + std::to_string((int)r));
            return VSGPU_ERR_CUDA;
        }
    }
#define VS_LAUNCH_TMA(QPV)                                                                                             \
    do {                                                                                                               \
        auto kern = scan_tma_kernel<ET, QPV, L2, TOPK>;                                                                \
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));             \
        kern<<<grid ? grid : (unsigned)std::min<size_t>((size_t)sm_count(s->device), (a.n + TILE_ROWS - 1) / TILE_ROWS), \
               SCAN_THREADS, smem_bytes, s->stream>>>(map, a);                                                              \
    } while (0)
    if (a.nq <= 4) VS_LAUNCH_TMA(2);
    else if (a.nq <= 8) VS_LAUNCH_TMA(4);
    else {
        if constexpr (TOPK) {
            set_error("fused scan: at most 8 queries per pass");
            return VSGPU_ERR_ARG;
        } else {
            VS_LAUNCH_TMA(8);
        }
    }
#undef VS_LAUNCH_TMA
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

} // namespace

bool tma_scan_supported(const vsgpu_store *s) {
    const ChainPlan &p = s->plan;
    if (p.kind != CK_LANES || p.G != 32 || p.prefix != 0) return false;
    if (s->type != VSGPU_FLOAT32 && s->type != VSGPU_FLOAT16) return false;
    return s->dim % 32 == 0 && s->dim <= 4096 && encode_fn() != nullptr;
}

// stage geometry of the staged scan for `qp` query pairs; `extra` bytes of shared memory are taken by the caller
static bool scan_geometry(const vsgpu_store *s, int qp, size_t extra, int *kc_out, int *nst_out, size_t *smem_out) {
    const int dim = (int)s->dim;
    const size_t esz = s->elem;
    const size_t q_bytes = (((size_t)qp * dim * sizeof(float2) + 127) / 128) * 128;
    int dev_smem = 0;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device) != cudaSuccess || dev_smem <= 0)
        dev_smem = 227 * 1024;
    if ((size_t)dev_smem < 128 + q_bytes + 1024 + extra + 3 * (size_t)TILE_ROWS * 32 * esz) return false;
    const size_t budget = (size_t)dev_smem - 128 - q_bytes - 1024 - extra;
    int kc = 0, nst = 0;
    for (int c = std::min(dim, 256); c >= 32; c -= 32) {
        if (dim % c) continue;
        const size_t stage = (size_t)TILE_ROWS * c * esz;
        const int n = (int)std::min<size_t>(budget / stage, MAX_STAGES);
        if (n >= 3) {
            kc = c;
            nst = n;
            break;
        }
    }
    if (!kc) return false;
    nst = std::min(nst, 6);
    *kc_out = kc;
    *nst_out = nst;
    *smem_out = 128 + q_bytes + extra + (size_t)nst * TILE_ROWS * kc * esz;
    return true;
}

// Scan + selection in one pass (north star: "only K scores ever leave the SM"): the k best (score, id) of up to 8 queries,
// ascending, into out_* ([nq][out_ld]). Two launches: the staged scan with a per-CTA running top-k, and the merge.
bool fused_topk_supported(const vsgpu_store *s, size_t nq, size_t k) {
    static const bool off = getenv("VSGPU_NO_FUSED_TOPK") != nullptr; // A/B switch
    return !off && tma_scan_supported(s) && nq >= 1 && nq <= 8 && k >= 1 && k <= (size_t)FUSED_K_MAX && k <= s->count;
}

int launch_fused_topk(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, size_t k, size_t out_ld, uint32_t *out_ids,
                      void *out_scores, uint64_t *out_labels) {
    const int qp = nq <= 4 ? 2 : 4;
    const size_t extra = (size_t)(2 * qp) * FUSED_CAP * 8 + 128;
    int kc = 0, nst = 0;
    size_t smem_bytes = 0;
    if (!scan_geometry(s, qp, extra, &kc, &nst, &smem_bytes)) {
        set_error("launch_fused_topk: dimension too large for the staged scan");
        return VSGPU_ERR_ARG;
    }
    const size_t ntiles = (s->count + TILE_ROWS - 1) / TILE_ROWS;
    // short stores: fewer CTAs than SMs would each hold a whole list of k; keep >= 2 k rows per CTA so the lists stay useful
    unsigned grid = (unsigned)std::min<size_t>((size_t)sm_count(s->device), ntiles);
    grid = (unsigned)std::max<size_t>(1, std::min<size_t>(grid, s->count / std::max<size_t>(2 * k, 64) + 1));
    VS_TRY(ensure_scratch(s, s->sel_state, nq * (size_t)grid * k * sizeof(uint2) + 256));
    ScanTmaArgs a{};
    a.rows = s->rows;
    a.row_stride = s->row_stride;
    a.n = s->count;
    a.dim = (int)s->dim;
    a.kc = kc;
    a.nstages = nst;
    a.q = (const uint8_t *)q_dev;
    a.q_stride = q_stride;
    a.nq = (int)nq;
    a.k = (int)k;
    a.part = (uint2 *)s->sel_state.ptr;
    const bool l2 = s->plan.is_l2;
    int rc;
    if (s->type == VSGPU_FLOAT32) rc = l2 ? launch_t<float, true, true>(s, a, smem_bytes, grid) : launch_t<float, false, true>(s, a, smem_bytes, grid);
    else rc = l2 ? launch_t<__half, true, true>(s, a, smem_bytes, grid) : launch_t<__half, false, true>(s, a, smem_bytes, grid);
    VS_TRY(rc);
    fused_merge_kernel<256><<<(unsigned)nq, 256, 0, s->stream>>>(a.part, (int)grid, (int)k, s->labels, (uint32_t)out_ld, out_ids,
                                                                (float *)out_scores, out_labels);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

// scores[q * ld + id] for up to 16 queries (raw blobs on the device)
int launch_tma_scan(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, void *scores, size_t ld) {
    if (nq == 0 || s->count == 0) return VSGPU_OK;
    if (nq > 16) {
        set_error("launch_tma_scan: at most 16 queries per pass");
        return VSGPU_ERR_ARG;
    }
    const int dim = (int)s->dim;
    const size_t esz = s->elem;
    const int qp = nq <= 4 ? 2 : (nq <= 8 ? 4 : 8);
    const size_t q_bytes = (((size_t)qp * dim * sizeof(float2) + 127) / 128) * 128;
    int dev_smem = 0;
    if (cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device) != cudaSuccess || dev_smem <= 0)
        dev_smem = 227 * 1024;
    const size_t budget = (size_t)dev_smem - 128 - q_bytes - 1024;
    // widest row segment that still leaves >= 3 stages in flight
    int kc = 0, nst = 0;
    for (int c = std::min(dim, 256); c >= 32; c -= 32) {
        if (dim % c) continue;
        const size_t stage = (size_t)TILE_ROWS * c * esz;
        const int n = (int)std::min<size_t>(budget / stage, MAX_STAGES);
        if (n >= 3) {
            kc = c;
            nst = n;
            break;
        }
    }
    if (!kc) {
        set_error("launch_tma_scan: dimension too large for the staged scan");
        return VSGPU_ERR_ARG;
    }
    nst = std::min(nst, 6);
    ScanTmaArgs a{};
    a.rows = s->rows;
    a.row_stride = s->row_stride;
    a.n = s->count;
    a.dim = dim;
    a.kc = kc;
    a.nstages = nst;
    a.q = (const uint8_t *)q_dev;
    a.q_stride = q_stride;
    a.nq = (int)nq;
    a.scores = (float *)scores;
    a.ld = ld;
    const size_t smem_bytes = 128 + q_bytes + (size_t)nst * TILE_ROWS * kc * esz;
    const bool l2 = s->plan.is_l2;
    if (s->type == VSGPU_FLOAT32) return l2 ? launch_t<float, true>(s, a, smem_bytes) : launch_t<float, false>(s, a, smem_bytes);
    return l2 ? launch_t<__half, true>(s, a, smem_bytes) : launch_t<__half, false>(s, a, smem_bytes);
}

} // namespace vsgpu
