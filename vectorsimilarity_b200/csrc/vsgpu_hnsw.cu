// HNSW on the device: batched graph traversal (one CTA per query) and a sequential, reference-order
// graph builder (one CTA per insertion). What each kernel re-states (paths relative to
// /root/reference/src/VecSim/algorithms/hnsw):
//   hnsw_search_kernel   HNSWIndex::topKQuery hnsw.h:2037-2084 = searchBottomLayerEP :1967-1981 +
//                        greedySearchLevel :1210-1258 + searchBottomLayer_WithTimeout :1983-2035 +
//                        processCandidate :530-613
//   hnsw_range_kernel    rangeQuery :2152-2200, searchRangeBottomLayer_WithTimeout :2086-2150,
//                        processCandidate_RangeSearch :615-680
//   hnsw_insert_kernel   appendVector/indexVector :1930-1960, insertElementToGraph :1567-1602,
//                        searchLayer :682-721, mutuallyConnectNewElement :870-941,
//                        getNeighborsByHeuristic2 :725-799, revisitNeighborConnections :801-868
//
// B200 mapping. A traversal is a chain of dependent random reads, so the unit of parallelism is the
// query: one CTA owns one query, its heaps live in shared memory, and each hop fans the <= M0
// neighbour distance evaluations out over the CTA's warps (all row loads of a hop in flight at
// once), with the reference's bit-exact operation order per distance (vsgpu_dist.cuh). Heap
// admission runs in link order on one thread, which is exactly the reference's sequential rule, so
// ids/scores are identical, not just close. Hundreds of queries are resident at once; visited sets
// are per-query bitmaps in HBM/L2.
#include "vsgpu_dist.cuh"
#include <algorithm>
#include <cstdlib>
#include <limits>
#include <vector>

namespace vsgpu {

static constexpr uint32_t INV = 0xffffffffu;
template <typename DT> __device__ __forceinline__ DT dt_max();
template <> __device__ __forceinline__ float dt_max<float>() { return 3.402823466e+38f; }
template <> __device__ __forceinline__ double dt_max<double>() { return 1.7976931348623157e+308; }
template <typename DT> __device__ __forceinline__ DT dt_nan();
template <> __device__ __forceinline__ float dt_nan<float>() { return __int_as_float(0x7fc00000); }
template <> __device__ __forceinline__ double dt_nan<double>() { return __longlong_as_double(0x7ff8000000000000ll); }
static constexpr int HNSW_THREADS = 256;
static constexpr int HEUR_CHUNK = 16; // candidates examined per speculative round of the heuristic

struct KCtx {
    const uint8_t *rows;
    size_t row_stride;
    int type, metric;
    ChainPlan plan;
    const float *norms;
    int chunks; // row_stride / 16 (integer types)
};

struct GraphDev {
    uint32_t *l0;           // [capacity][M0 + 1]: count, links
    uint32_t *up;           // [records][M + 1]
    const uint32_t *up_off; // [capacity]: first upper record of a node
    const uint32_t *levels; // [capacity]: top level
    const uint8_t *flags;   // [capacity]: bit0 = marked deleted
    int *state;             // [0] entry point (-1 = none), [1] max level
    int M, M0;
};

__device__ __forceinline__ uint32_t *links_of(const GraphDev &g, uint32_t node, int level) {
    return level == 0 ? g.l0 + (size_t)node * (g.M0 + 1) : g.up + ((size_t)g.up_off[node] + (level - 1)) * (g.M + 1);
}
__device__ __forceinline__ bool is_deleted(const GraphDev &g, uint32_t node) { return g.flags[node] & 1; }

// ------------------------------------------------------------------------------------------------
// Distance policies: P::dists<PIVOT> evaluates RU (row, other) pairs per thread group of G lanes;
// `other` is the pivot staged in shared memory (the query) or a second stored row.
// element load with the storage type fixed at compile time: no branch between the loads of a hop, so the
// compiler issues them back to back (with a run-time type switch every load waited for the previous one:
// 16 dependent L2 round trips per distance evaluation, measured 6.7 k cycles per hop)
template <int ST, typename DT> __device__ __forceinline__ DT load_st(const uint8_t *row, int e) {
    if constexpr (ST == VSGPU_FLOAT64) return (DT)__ldg(reinterpret_cast<const double *>(row) + e);
    else if constexpr (ST == VSGPU_FLOAT32) return (DT)__ldg(reinterpret_cast<const float *>(row) + e);
    else if constexpr (ST == VSGPU_BFLOAT16) return (DT)__uint_as_float((unsigned)__ldg(reinterpret_cast<const unsigned short *>(row) + e) << 16);
    else return (DT)__half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short *>(row) + e)));
}

template <typename CT_, int G_, bool FTZ, bool L2, int ST> struct PolChain {
    using DT = CT_;
    static constexpr int G = G_;
    static constexpr int RU = sizeof(CT_) == 8 ? 2 : 4;
    static size_t pivot_bytes(const vsgpu_store *s) { return (size_t)s->plan.S * G * sizeof(DT); }
    __device__ static void load_pivot(const KCtx &k, void *pv_, const uint8_t *src, float) {
        DT *pv = (DT *)pv_;
        const int total = k.plan.S * G;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int e = chain_elem(k.plan, i % G, i / G);
            pv[i] = e < 0 ? DT(0) : load_st<ST, DT>(src, e);
        }
    }
    template <bool PIVOT>
    __device__ static void dists(const KCtx &k, const void *pv_, const uint32_t (&a)[RU], const uint32_t (&b)[RU], int c,
                                 DT (&out)[RU]) {
        const DT *pv = (const DT *)pv_ + c;
        const uint8_t *ra[RU], *rb[RU];
#pragma unroll
        for (int r = 0; r < RU; r++) {
            ra[r] = k.rows + (size_t)(a[r] == INV ? 0 : a[r]) * k.row_stride;
            rb[r] = k.rows + (size_t)(b[r] == INV ? 0 : b[r]) * k.row_stride;
        }
        DT acc[RU];
#pragma unroll
        for (int r = 0; r < RU; r++) acc[r] = DT(0);
        const bool fast = k.plan.kind == CK_LANES && k.plan.prefix == 0;
        const int S = k.plan.S;
        if (fast) {
            // no residual: chain c reads element G*s + c — straight-line loads, 4 steps x RU rows in flight
#pragma unroll 4
            for (int s = 0; s < S; s++) {
                const int e = G * s + c;
#pragma unroll
                for (int r = 0; r < RU; r++) {
                    const DT x = load_st<ST, DT>(ra[r], e);
                    DT y;
                    if constexpr (PIVOT) y = pv[s * G];
                    else y = load_st<ST, DT>(rb[r], e);
                    if constexpr (L2) {
                        const DT d = sub_rn(x, y);
                        acc[r] = fma_step<FTZ>(d, d, acc[r]);
                    } else {
                        acc[r] = fma_step<FTZ>(x, y, acc[r]);
                    }
                }
            }
        } else {
            for (int s = 0; s < S; s++) {
                const int e = chain_elem(k.plan, c, s);
                const int ee = e < 0 ? 0 : e; // padded slots: load element 0, multiply by zero below is NOT exact for inf/nan -> select
                DT x[RU], y[RU];
#pragma unroll
                for (int r = 0; r < RU; r++) {
                    x[r] = load_st<ST, DT>(ra[r], ee);
                    if constexpr (!PIVOT) y[r] = load_st<ST, DT>(rb[r], ee);
                }
#pragma unroll
                for (int r = 0; r < RU; r++) {
                    const DT xv = e < 0 ? DT(0) : x[r];
                    DT yv;
                    if constexpr (PIVOT) yv = pv[s * G];
                    else yv = e < 0 ? DT(0) : y[r];
                    if constexpr (L2) {
                        const DT d = sub_rn(xv, yv);
                        acc[r] = fma_step<FTZ>(d, d, acc[r]);
                    } else {
                        acc[r] = fma_step<FTZ>(xv, yv, acc[r]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RU; r++) {
            DT v = butterfly<DT, G>(acc[r]);
            if (!L2) v = sub_rn(DT(1), v);
            out[r] = v;
        }
    }
    // NR rows against the staged pivot with every load of a 4-step slice in flight at once (plans without a residual:
    // chain c reads element G*s + c). Same chains, same order, same butterfly as dists<true>: bit-identical results.
    // Used by the warp-per-query traversal, where one warp has to hide the row latency on its own.
    template <int NR>
    __device__ static void dists_fast(const KCtx &k, const void *pv_, const uint32_t (&a)[NR], int c, DT (&out)[NR]) {
        const DT *pv = (const DT *)pv_ + c;
        const uint8_t *ra[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) ra[r] = k.rows + (size_t)(a[r] == INV ? 0 : a[r]) * k.row_stride;
        DT acc[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) acc[r] = DT(0);
        const int S = k.plan.S;
        for (int s0 = 0; s0 < S; s0 += 4) {
            DT x[NR][4];
#pragma unroll
            for (int t = 0; t < 4; t++)
#pragma unroll
                for (int r = 0; r < NR; r++) x[r][t] = s0 + t < S ? load_st<ST, DT>(ra[r], G * (s0 + t) + c) : DT(0);
#pragma unroll
            for (int t = 0; t < 4; t++) {
                if (s0 + t < S) {
                    const DT y = pv[(s0 + t) * G];
#pragma unroll
                    for (int r = 0; r < NR; r++) {
                        if constexpr (L2) {
                            const DT d = sub_rn(x[r][t], y);
                            acc[r] = fma_step<FTZ>(d, d, acc[r]);
                        } else {
                            acc[r] = fma_step<FTZ>(x[r][t], y, acc[r]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < NR; r++) {
            DT v = butterfly<DT, G>(acc[r]);
            if (!L2) v = sub_rn(DT(1), v);
            out[r] = v;
        }
    }
    static constexpr bool HAS_FAST = true;
    // one thread evaluates a whole (row, row) pair: the G chain sums live in registers and are folded
    // with the same pairing tree as the warp butterfly (acc[c] + acc[c + w], w = G/2 .. 1). Used where
    // many independent pairs are wanted at once (neighbour-selection heuristic), so a thread per pair
    // gives far more loads in flight than a warp per pair.
    __device__ static DT dist_thread(const KCtx &k, uint32_t a, uint32_t b) {
        const uint8_t *ra = k.rows + (size_t)a * k.row_stride, *rb = k.rows + (size_t)b * k.row_stride;
        DT acc[G];
#pragma unroll
        for (int c = 0; c < G; c++) acc[c] = DT(0);
        const bool fast = k.plan.kind == CK_LANES && k.plan.prefix == 0;
        const int S = k.plan.S;
        if (ST == VSGPU_FLOAT32 && fast) {
            for (int s = 0; s < S; s++) {
#pragma unroll
                for (int c4 = 0; c4 < G / 4; c4++) {
                    const float4 x = __ldg(reinterpret_cast<const float4 *>(ra) + s * (G / 4) + c4);
                    const float4 y = __ldg(reinterpret_cast<const float4 *>(rb) + s * (G / 4) + c4);
                    const float xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        if constexpr (L2) {
                            const DT d = sub_rn((DT)xs[t], (DT)ys[t]);
                            acc[4 * c4 + t] = fma_step<FTZ>(d, d, acc[4 * c4 + t]);
                        } else {
                            acc[4 * c4 + t] = fma_step<FTZ>((DT)xs[t], (DT)ys[t], acc[4 * c4 + t]);
                        }
                    }
                }
            }
        } else {
            for (int s = 0; s < S; s++) {
#pragma unroll
                for (int c = 0; c < G; c++) {
                    const int e = fast ? G * s + c : chain_elem(k.plan, c, s);
                    const DT x = e < 0 ? DT(0) : load_st<ST, DT>(ra, e);
                    const DT y = e < 0 ? DT(0) : load_st<ST, DT>(rb, e);
                    if constexpr (L2) {
                        const DT d = sub_rn(x, y);
                        acc[c] = fma_step<FTZ>(d, d, acc[c]);
                    } else {
                        acc[c] = fma_step<FTZ>(x, y, acc[c]);
                    }
                }
            }
        }
#pragma unroll
        for (int w = G / 2; w >= 1; w >>= 1)
#pragma unroll
            for (int c = 0; c < G / 2; c++)
                if (c < w) acc[c] = add_rn(acc[c], acc[c + w]);
        return L2 ? acc[0] : sub_rn(DT(1), acc[0]);
    }
};

struct IntPivotTail {
    long long qq;
    float qn;
    int pad;
};
template <bool U> struct PolInt {
    using DT = float;
    static constexpr int G = 8;
    static constexpr int RU = 2;
    static constexpr bool HAS_FAST = false;
    static size_t pivot_bytes(const vsgpu_store *s) { return s->row_stride + sizeof(IntPivotTail); }
    // src: dim bytes zero padded to row_stride (a stored row, or a staged query)
    __device__ static void load_pivot(const KCtx &k, void *pv_, const uint8_t *src, float norm) {
        uint4 *pv = (uint4 *)pv_;
        for (int i = threadIdx.x; i < k.chunks; i += blockDim.x) pv[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
        if (threadIdx.x == 0) {
            long long t = 0;
            for (int i = 0; i < k.chunks; i++) {
                const uint4 v = __ldg(reinterpret_cast<const uint4 *>(src) + i);
                t += dot4<U>(v.x, v.x, 0) + dot4<U>(v.y, v.y, 0) + dot4<U>(v.z, v.z, 0) + dot4<U>(v.w, v.w, 0);
            }
            IntPivotTail *tail = reinterpret_cast<IntPivotTail *>((uint8_t *)pv_ + k.row_stride);
            tail->qq = t;
            tail->qn = norm;
        }
    }
    template <bool PIVOT>
    __device__ static void dists(const KCtx &k, const void *pv_, const uint32_t (&a)[RU], const uint32_t (&b)[RU], int c,
                                 DT (&out)[RU]) {
        const uint4 *pv = (const uint4 *)pv_;
        const IntPivotTail *tail = reinterpret_cast<const IntPivotTail *>((const uint8_t *)pv_ + k.row_stride);
#pragma unroll
        for (int r = 0; r < RU; r++) {
            const uint32_t ia = a[r] == INV ? 0 : a[r], ib = b[r] == INV ? 0 : b[r];
            const uint4 *ra = reinterpret_cast<const uint4 *>(k.rows + (size_t)ia * k.row_stride);
            const uint4 *rb = reinterpret_cast<const uint4 *>(k.rows + (size_t)ib * k.row_stride);
            long long dot = 0, aa = 0, bb = 0;
            for (int ch = c; ch < k.chunks; ch += G) {
                const uint4 x = __ldg(ra + ch);
                uint4 y;
                if constexpr (PIVOT) y = pv[ch];
                else y = __ldg(rb + ch);
                dot += dot4<U>(x.x, y.x, 0) + dot4<U>(x.y, y.y, 0) + dot4<U>(x.z, y.z, 0) + dot4<U>(x.w, y.w, 0);
                aa += dot4<U>(x.x, x.x, 0) + dot4<U>(x.y, x.y, 0) + dot4<U>(x.z, x.z, 0) + dot4<U>(x.w, x.w, 0);
                if constexpr (!PIVOT)
                    bb += dot4<U>(y.x, y.x, 0) + dot4<U>(y.y, y.y, 0) + dot4<U>(y.z, y.z, 0) + dot4<U>(y.w, y.w, 0);
            }
#pragma unroll
            for (int w = G / 2; w >= 1; w >>= 1) {
                dot += __shfl_xor_sync(0xffffffffu, dot, w);
                aa += __shfl_xor_sync(0xffffffffu, aa, w);
                if constexpr (!PIVOT) bb += __shfl_xor_sync(0xffffffffu, bb, w);
            }
            const float rn = k.norms ? k.norms[ia] : 0.f;
            const float qn = PIVOT ? tail->qn : (k.norms ? k.norms[ib] : 0.f);
            out[r] = int_score(k.metric, dot, aa, PIVOT ? tail->qq : bb, rn, qn);
        }
    }
    __device__ static DT dist_thread(const KCtx &k, uint32_t a, uint32_t b) {
        const uint4 *ra = reinterpret_cast<const uint4 *>(k.rows + (size_t)a * k.row_stride);
        const uint4 *rb = reinterpret_cast<const uint4 *>(k.rows + (size_t)b * k.row_stride);
        long long dot = 0, aa = 0, bb = 0;
        for (int ch = 0; ch < k.chunks; ch++) {
            const uint4 x = __ldg(ra + ch), y = __ldg(rb + ch);
            dot += dot4<U>(x.x, y.x, 0) + dot4<U>(x.y, y.y, 0) + dot4<U>(x.z, y.z, 0) + dot4<U>(x.w, y.w, 0);
            aa += dot4<U>(x.x, x.x, 0) + dot4<U>(x.y, x.y, 0) + dot4<U>(x.z, x.z, 0) + dot4<U>(x.w, x.w, 0);
            bb += dot4<U>(y.x, y.x, 0) + dot4<U>(y.y, y.y, 0) + dot4<U>(y.z, y.z, 0) + dot4<U>(y.w, y.w, 0);
        }
        return int_score(k.metric, dot, aa, bb, k.norms ? k.norms[a] : 0.f, k.norms ? k.norms[b] : 0.f);
    }
};

template <typename CT_> struct PolSeq {
    using DT = CT_;
    static constexpr int G = 1;
    static constexpr int RU = 1;
    static constexpr bool HAS_FAST = false;
    static size_t pivot_bytes(const vsgpu_store *s) { return s->dim * sizeof(DT); }
    __device__ static void load_pivot(const KCtx &k, void *pv_, const uint8_t *src, float) {
        DT *pv = (DT *)pv_;
        for (int i = threadIdx.x; i < k.plan.dim; i += blockDim.x) pv[i] = Loader<DT>::load(src, k.type, i);
    }
    template <bool PIVOT>
    __device__ static void dists(const KCtx &k, const void *pv_, const uint32_t (&a)[RU], const uint32_t (&b)[RU], int,
                                 DT (&out)[RU]) {
        const DT *pv = (const DT *)pv_;
        const uint8_t *ra = k.rows + (size_t)(a[0] == INV ? 0 : a[0]) * k.row_stride;
        const uint8_t *rb = k.rows + (size_t)(b[0] == INV ? 0 : b[0]) * k.row_stride;
        if constexpr (PIVOT) out[0] = seq_dist<DT>(ra, k.type, k.plan, [&](int e) { return pv[e]; });
        else out[0] = seq_dist<DT>(ra, k.type, k.plan, [&](int e) { return Loader<DT>::load(rb, k.type, e); });
    }
    __device__ static DT dist_thread(const KCtx &k, uint32_t a, uint32_t b) {
        const uint8_t *ra = k.rows + (size_t)a * k.row_stride, *rb = k.rows + (size_t)b * k.row_stride;
        return seq_dist<DT>(ra, k.type, k.plan, [&](int e) { return Loader<DT>::load(rb, k.type, e); });
    }
};

// out[j] = dist(pair j) for j < n; pair(j, &a, &b) names the rows (a == INV: skip). PIVOT: `b` is
// ignored and the staged pivot is the other operand. Called by every thread of the CTA.
template <class P, bool PIVOT, typename PairFn>
__device__ __forceinline__ void eval_dists(const KCtx &k, const void *pv, int n, typename P::DT *out, PairFn pair) {
    constexpr int GPW = 32 / P::G, RU = P::RU;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int c = lane % P::G, grp = lane / P::G;
    const int per_round = nw * GPW * RU;
    for (int base = 0; base < n; base += per_round) {
        if (base + warp * GPW * RU >= n) continue; // warp-uniform
        const int j0 = base + (warp * GPW + grp) * RU;
        uint32_t a[RU], b[RU];
#pragma unroll
        for (int r = 0; r < RU; r++) {
            a[r] = INV;
            b[r] = INV;
            if (j0 + r < n) pair(j0 + r, a[r], b[r]);
        }
        typename P::DT o[RU];
        P::template dists<PIVOT>(k, pv, a, b, c, o);
        if (c == 0) {
#pragma unroll
            for (int r = 0; r < RU; r++)
                if (a[r] != INV) out[j0 + r] = o[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// binary heaps driven by one thread; Less(a_d, a_id, b_d, b_id) is the reference's pair ordering
template <typename DT, typename Less>
__device__ __forceinline__ void heap_push(DT *hd, uint32_t *hid, int &n, DT d, uint32_t id, Less less) {
    int i = n++;
    while (i > 0) {
        const int p = (i - 1) >> 1;
        if (!less(hd[p], hid[p], d, id)) break;
        hd[i] = hd[p];
        hid[i] = hid[p];
        i = p;
    }
    hd[i] = d;
    hid[i] = id;
}
template <typename DT, typename Less> __device__ __forceinline__ void heap_pop(DT *hd, uint32_t *hid, int &n, Less less) {
    n--;
    if (n == 0) return;
    const DT d = hd[n];
    const uint32_t id = hid[n];
    int i = 0;
    for (;;) {
        int ch = 2 * i + 1;
        if (ch >= n) break;
        if (ch + 1 < n && less(hd[ch], hid[ch], hd[ch + 1], hid[ch + 1])) ch++;
        if (!less(d, id, hd[ch], hid[ch])) break;
        hd[i] = hd[ch];
        hid[i] = hid[ch];
        i = ch;
    }
    hd[i] = d;
    hid[i] = id;
}

// std::pair<DistType, key> operator< (top candidates: key = label for queries, id for the builder)
template <typename DT> struct TopLess {
    const uint64_t *labels; // nullptr: key = id
    __device__ __forceinline__ bool operator()(DT ad, uint32_t ai, DT bd, uint32_t bi) const {
        if (ad < bd) return true;
        if (bd < ad) return false;
        if (labels) return labels[ai] < labels[bi];
        return ai < bi;
    }
};
// candidate set: std::pair<DistType, idType>(-dist, id) under operator<
template <typename DT> struct CandLess {
    __device__ __forceinline__ bool operator()(DT ad, uint32_t ai, DT bd, uint32_t bi) const {
        if (-ad < -bd) return true;
        if (-bd < -ad) return false;
        return ai < bi;
    }
};

struct Visited {
    uint32_t *p;
    uint32_t tag; // 0: p is a bitmap (atomic), else p is a tag array
};
__device__ __forceinline__ bool test_and_set(const Visited &v, uint32_t id) {
    if (v.tag) {
        if (v.p[id] == v.tag) return true;
        v.p[id] = v.tag;
        return false;
    }
    const uint32_t bit = 1u << (id & 31);
    return (atomicOr(&v.p[id >> 5], bit) & bit) != 0;
}

// Per-CTA working set (pointers into shared memory, cand_* may be re-pointed to the HBM spill area)
template <typename DT> struct Work {
    void *pivot;
    uint32_t *nb_ids;
    uint8_t *nb_del; // deleted flag of each gathered neighbour (fetched with the links, off the serial path)
    DT *nb_dist;
    DT *top_d;
    uint32_t *top_id;
    DT *cand_d;
    uint32_t *cand_id;
    int cand_cap;
    DT *spill_d; // HBM spill for the candidate set (capacity = element count) or nullptr
    uint32_t *spill_id;
    int spill_cap;
    int *sc; // shared scalars
    DT *sdt; // shared DT scalars: [0] lowerBound / curDist
    // admission log (builder): the first `adm_cap` pushes to the top heap, in order
    DT *adm_d;
    uint32_t *adm_id;
    int adm_cap;
    long long *prof; // thread 0: cycles per phase (gather, eval, admit, pop), or nullptr
    // read log (batched builder): every node whose link list this traversal read
    uint32_t *rlog;
    int rlog_cap;
    int *rlog_n; // shared counter; > cap = overflow
    bool no_regtop; // A/B switch (VSGPU_HNSW_NO_REGTOP): keep the result set in shared memory even for ef <= 64
};
enum { SC_TOPN = 0, SC_CANDN, SC_NBN, SC_STOP, SC_CUR, SC_STATUS, SC_ADMN, SC_AUX0, SC_AUX1, SC_AUX2, SC_AUX3, SC_COUNT = 16 };

// warp 0: links of `node` at `level` that were not visited yet, in link order -> w.nb_ids / nb_del, SC_NBN.
// The count and the link words are fetched together (the record is always fully allocated), the
// visited test, the deleted flag and an L2 prefetch of the row are issued back to back: one hop costs
// two dependent memory round trips before the distance evaluation instead of four.
template <typename DT>
__device__ __forceinline__ void gather_unvisited(const KCtx &k, const GraphDev &g, const Work<DT> &w, uint32_t node, int level,
                                                 const Visited *vis) {
    if (threadIdx.x >= 32) return;
    const uint32_t *rec = links_of(g, node, level);
    const int width = level == 0 ? g.M0 : g.M;
    int base = 0;
    int cnt = 0;
    for (int i0 = 0; i0 < width; i0 += 32) {
        const int i = i0 + threadIdx.x;
        uint32_t id = i < width ? rec[1 + i] : INV;
        if (i0 == 0) cnt = (int)rec[0];
        if (i0 >= cnt) break;
        bool take = false;
        uint8_t del = 0;
        if (i < cnt) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(k.rows + (size_t)id * k.row_stride));
            del = g.flags[id];
            take = vis ? !test_and_set(*vis, id) : true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (take) {
            const int pos = base + __popc(m & ((1u << threadIdx.x) - 1));
            w.nb_ids[pos] = id;
            w.nb_del[pos] = del & 1;
        }
        base += __popc(m);
    }
    if (threadIdx.x == 0) w.sc[SC_NBN] = base;
}

// greedySearchLevel (hnsw.h:1210-1258). cur/curDist in w.sc[SC_CUR]/w.sdt[0]. `track_deleted`: the
// builder's variant that hands the best non-deleted node to the next level.
template <class P>
__device__ void greedy_level(const KCtx &k, const GraphDev &g, const Work<typename P::DT> &w, int level, bool track_deleted,
                             unsigned long long &evals) {
    using DT = typename P::DT;
    if (threadIdx.x == 0) w.sc[SC_AUX0] = w.sc[SC_CUR]; // bestNonDeletedCand
    for (;;) {
        __syncthreads();
        const uint32_t cur = (uint32_t)w.sc[SC_CUR];
        if (w.rlog && threadIdx.x == 0) {
            const int p = (*w.rlog_n)++;
            if (p < w.rlog_cap) w.rlog[p] = cur;
        }
        gather_unvisited<DT>(k, g, w, cur, level, nullptr);
        __syncthreads();
        const int n = w.sc[SC_NBN];
        eval_dists<P, true>(k, w.pivot, n, w.nb_dist, [&](int j, uint32_t &a, uint32_t &) { a = w.nb_ids[j]; });
        __syncthreads();
        if (threadIdx.x == 0) {
            bool changed = false;
            DT best = w.sdt[0];
            for (int j = 0; j < n; j++) {
                if (w.nb_dist[j] < best) {
                    best = w.nb_dist[j];
                    w.sc[SC_CUR] = (int)w.nb_ids[j];
                    changed = true;
                    if (track_deleted && !w.nb_del[j]) w.sc[SC_AUX0] = (int)w.nb_ids[j];
                }
            }
            w.sdt[0] = best;
            w.sc[SC_STOP] = changed ? 0 : 1;
            evals += n;
        }
        __syncthreads();
        if (w.sc[SC_STOP]) break;
    }
    if (track_deleted && threadIdx.x == 0) w.sc[SC_CUR] = w.sc[SC_AUX0];
    __syncthreads();
}

// ---- warp-cooperative sorted arrays (warp 0 only) ----
// Both traversal queues are kept sorted ascending under the reference's pair order, so the element a
// std::priority_queue would expose as top() is the LAST one: pop = n--. An insertion counts the
// elements ordered before the new one (strided over the lanes), then shifts the tail by one slot,
// 32 elements per step — a few hundred cycles where a one-thread binary heap in shared memory
// costs thousands (every sift level is a dependent LDS). Any container with the same total order
// pops the same sequence, so results are unchanged.
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    return v;
}
template <typename DT, typename Less>
__device__ __forceinline__ void sorted_insert(DT *ad, uint32_t *aid, int &n, DT d, uint32_t id, Less less) {
    const int lane = threadIdx.x & 31;
    // position = number of elements ordered before the new one: 32-ary search over the sorted array
    int lo = 0, hi_s = n;
    while (hi_s - lo > 32) {
        const int step = (hi_s - lo + 31) / 32;
        const int idx = lo + lane * step;
        const bool pred = idx < hi_s && less(ad[idx], aid[idx], d, id);
        const int t = __popc(__ballot_sync(0xffffffffu, pred)); // predicates are monotone: t leading trues
        if (t == 0) {
            hi_s = lo;
            break;
        }
        const int nlo = lo + (t - 1) * step + 1;
        hi_s = min(lo + t * step, hi_s);
        lo = nlo;
    }
    int cnt = lo;
    if (hi_s > lo) {
        const int idx = lo + lane;
        const bool pred = idx < hi_s && less(ad[idx], aid[idx], d, id);
        cnt = lo + __popc(__ballot_sync(0xffffffffu, pred));
    }
    for (int hi = n - 1; hi >= cnt; hi -= 32) {
        const int i = hi - lane;
        const bool act = i >= cnt;
        DT td = DT(0);
        uint32_t ti = 0;
        if (act) {
            td = ad[i];
            ti = aid[i];
        }
        __syncwarp();
        if (act) {
            ad[i + 1] = td;
            aid[i + 1] = ti;
        }
        __syncwarp();
    }
    if (lane == 0) {
        ad[cnt] = d;
        aid[cnt] = id;
    }
    __syncwarp();
    n++;
}
// drop the first t elements
template <typename DT> __device__ __forceinline__ void sorted_drop_prefix(DT *ad, uint32_t *aid, int &n, int t) {
    const int lane = threadIdx.x & 31;
    if (t <= 0) return;
    for (int lo = t; lo < n; lo += 32) {
        const int i = lo + lane;
        const bool act = i < n;
        DT td = DT(0);
        uint32_t ti = 0;
        if (act) {
            td = ad[i];
            ti = aid[i];
        }
        __syncwarp();
        if (act) {
            ad[i - t] = td;
            aid[i - t] = ti;
        }
        __syncwarp();
    }
    n -= t;
}

// ---- candidate set: an unordered bag (warp 0) ----
// Adding is one store; the element a priority queue would pop — the maximum under pair(-dist, id) — is
// found by a strided scan + warp arg-max when it is needed, once per hop. (A sorted array cost ~2.5 k
// cycles per admission at efConstruction-sized sets; the hop's single pop costs less than one of those.)
template <typename DT> __device__ __forceinline__ int cand_best(const DT *cd, const uint32_t *cid, int n, DT &bd, uint32_t &bid) {
    CandLess<DT> cl;
    const int lane = threadIdx.x & 31;
    int bi = -1;
    for (int i = lane; i < n; i += 32) {
        const DT d = cd[i];
        const uint32_t id = cid[i];
        if (bi < 0 || cl(bd, bid, d, id)) {
            bd = d;
            bid = id;
            bi = i;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const DT od = __shfl_xor_sync(0xffffffffu, bd, m);
        const uint32_t oid = __shfl_xor_sync(0xffffffffu, bid, m);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
        if (oi >= 0 && (bi < 0 || cl(bd, bid, od, oid))) {
            bd = od;
            bid = oid;
            bi = oi;
        }
    }
    return bi;
}
template <typename DT> __device__ __forceinline__ void cand_remove(DT *cd, uint32_t *cid, int &n, int idx) {
    if ((threadIdx.x & 31) == 0) {
        cd[idx] = cd[n - 1];
        cid[idx] = cid[n - 1];
    }
    __syncwarp();
    n--;
}
// add; on overflow first drop entries that can no longer be expanded, then spill to HBM. false = out of room.
template <typename DT>
__device__ __forceinline__ bool cand_insert(Work<DT> &w, int &cand_n, DT d, uint32_t id, bool can_prune, DT lower) {
    const int lane = threadIdx.x & 31;
    if (cand_n >= w.cand_cap) {
        if (can_prune) {
            // entries farther than the current bound are never expanded: the bound only shrinks once the
            // result set is full, and the stop rule fires before they are reached
            int out = 0;
            for (int base = 0; base < cand_n; base += 32) {
                const int i = base + lane;
                DT td = DT(0);
                uint32_t ti = 0;
                bool keep = false;
                if (i < cand_n) {
                    td = w.cand_d[i];
                    ti = w.cand_id[i];
                    keep = !(td > lower);
                }
                const unsigned m = __ballot_sync(0xffffffffu, keep);
                __syncwarp();
                if (keep) {
                    const int pos = out + __popc(m & ((1u << lane) - 1));
                    w.cand_d[pos] = td;
                    w.cand_id[pos] = ti;
                }
                __syncwarp();
                out += __popc(m);
            }
            cand_n = out;
        }
        if (cand_n >= w.cand_cap) {
            if (!w.spill_d || w.cand_d == w.spill_d) return false;
            for (int i = lane; i < cand_n; i += 32) {
                w.spill_d[i] = w.cand_d[i];
                w.spill_id[i] = w.cand_id[i];
            }
            __syncwarp();
            w.cand_d = w.spill_d;
            w.cand_id = w.spill_id;
            w.cand_cap = w.spill_cap;
            if (cand_n >= w.cand_cap) return false;
        }
    }
    if (lane == 0) {
        w.cand_d[cand_n] = d;
        w.cand_id[cand_n] = id;
    }
    __syncwarp();
    cand_n++;
    return true;
}

// searchLayer / searchBottomLayer (hnsw.h:682-721, 1983-2035) from entry w.sc[SC_CUR]. Result set in
// w.top_* (SC_TOPN entries, ascending under TopLess). labels != nullptr: query flavour (keyed by
// label); else builder flavour (keyed by id, admissions logged). Status through w.sc[SC_STATUS].
// `multi` (HNSWIndex_Multi, hnsw_multi.h:62-71,103-106): the result set is an updatable_max_heap keyed by LABEL
// (utils/updatable_heap.h:66-111) — a label already in the set only ever improves its score, it never takes a second slot.
template <class P>
__device__ void search_layer(const KCtx &k, const GraphDev &g, Work<typename P::DT> &w, int level, int ef,
                             const uint64_t *labels, const Visited &vis, unsigned long long &evals,
                             unsigned long long &hops, bool multi = false) {
    using DT = typename P::DT;
    TopLess<DT> tl{labels};
    CandLess<DT> cl;
    const bool warp0 = threadIdx.x < 32;
    const int lane = threadIdx.x & 31;
    int top_n = 0, cand_n = 0, adm_n = 0; // warp 0's (uniform) copies
    DT lower = DT(0);
    // ef <= 64: warp 0 keeps the result set in registers, sorted under `tl`: lane L holds entries L and L + 32. An insertion
    // is two ballots and six shuffles instead of a 32-ary search plus a shift through shared memory (~250 cycles each, a
    // third of a hop); the arrays in shared memory are written once, when the search ends. A 65th entry falls off the end,
    // which is exactly "insert, then drop the largest" of a full set.
    const bool regtop = ef <= 64 && !w.no_regtop;
    DT t0d = DT(0), t1d = DT(0);
    uint32_t t0i = 0, t1i = 0;
    // multi, register-resident set: position of the entry that carries `lab`, or -1 (labels are fetched per probe: the
    // probe is one global load per lane, off the critical path of single-value searches, which never call this)
    auto reg_find_label = [&](uint64_t lab) -> int {
        const bool p0 = lane < top_n && labels[t0i] == lab;
        const bool p1 = lane + 32 < top_n && labels[t1i] == lab;
        const unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
        return b0 ? __ffs(b0) - 1 : (b1 ? 32 + __ffs(b1) - 1 : -1);
    };
    auto reg_dist_at = [&](int p) -> DT {
        const DT a = __shfl_sync(0xffffffffu, t0d, p & 31), b = __shfl_sync(0xffffffffu, t1d, p & 31);
        return p < 32 ? a : b;
    };
    auto reg_remove = [&](int p) { // entries after p move down by one
        const DT d0 = __shfl_down_sync(0xffffffffu, t0d, 1), d1 = __shfl_down_sync(0xffffffffu, t1d, 1);
        const uint32_t i0 = __shfl_down_sync(0xffffffffu, t0i, 1), i1 = __shfl_down_sync(0xffffffffu, t1i, 1);
        const DT f1d = __shfl_sync(0xffffffffu, t1d, 0); // entry 32 becomes entry 31
        const uint32_t f1i = __shfl_sync(0xffffffffu, t1i, 0);
        if (lane >= p) {
            t0d = lane == 31 ? f1d : d0;
            t0i = lane == 31 ? f1i : i0;
        }
        if (lane + 32 >= p) {
            t1d = d1;
            t1i = i1;
        }
        top_n--;
    };
    auto reg_insert = [&](DT d, uint32_t id) {
        const bool p0 = lane < top_n && tl(t0d, t0i, d, id);
        const bool p1 = lane + 32 < top_n && tl(t1d, t1i, d, id);
        const int cnt = __popc(__ballot_sync(0xffffffffu, p0)) + __popc(__ballot_sync(0xffffffffu, p1));
        const DT u0d = __shfl_up_sync(0xffffffffu, t0d, 1);
        const uint32_t u0i = __shfl_up_sync(0xffffffffu, t0i, 1);
        DT u1d = __shfl_up_sync(0xffffffffu, t1d, 1);
        uint32_t u1i = __shfl_up_sync(0xffffffffu, t1i, 1);
        const DT l31d = __shfl_sync(0xffffffffu, t0d, 31);
        const uint32_t l31i = __shfl_sync(0xffffffffu, t0i, 31);
        if (lane == 0) { // entry 32's left neighbour is entry 31
            u1d = l31d;
            u1i = l31i;
        }
        if (lane == cnt) {
            t0d = d;
            t0i = id;
        } else if (lane > cnt) {
            t0d = u0d;
            t0i = u0i;
        }
        if (lane + 32 == cnt) {
            t1d = d;
            t1i = id;
        } else if (lane + 32 > cnt) {
            t1d = u1d;
            t1i = u1i;
        }
        top_n = min(top_n + 1, 64);
    };
    auto reg_last = [&]() -> DT { // distance of entry top_n - 1
        const int li = top_n - 1;
        const DT a = __shfl_sync(0xffffffffu, t0d, li & 31), b = __shfl_sync(0xffffffffu, t1d, li & 31);
        return li < 32 ? a : b;
    };
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ep = (uint32_t)w.sc[SC_CUR];
        w.nb_ids[0] = ep;
        w.sc[SC_STATUS] = 0;
    }
    __syncthreads();
    eval_dists<P, true>(k, w.pivot, 1, w.nb_dist, [&](int j, uint32_t &a, uint32_t &) { a = w.nb_ids[j]; });
    __syncthreads();
    if (warp0) {
        const uint32_t ep = w.nb_ids[0];
        if (lane == 0) evals += 1;
        if (!is_deleted(g, ep)) {
            lower = w.nb_dist[0];
            if (regtop) reg_insert(lower, ep);
            else sorted_insert(w.top_d, w.top_id, top_n, lower, ep, tl);
            if (lane == 0 && w.adm_cap > 0) {
                w.adm_d[0] = lower;
                w.adm_id[0] = ep;
            }
            adm_n = 1;
        } else {
            lower = dt_max<DT>();
        }
        cand_insert(w, cand_n, lower, ep, false, lower);
        if (lane == 0) test_and_set(vis, ep);
    }
    for (;;) {
        if (warp0) {
            int stop = 0;
            DT bd = DT(0);
            uint32_t bid = 0;
            const int bi = cand_n ? cand_best(w.cand_d, w.cand_id, cand_n, bd, bid) : -1;
            if (bi < 0) stop = 1;
            else if (bd > lower && top_n >= ef) stop = 1;
            else {
                if (lane == 0) {
                    w.sc[SC_CUR] = (int)bid;
                    hops++;
                    if (w.rlog) {
                        const int p = (*w.rlog_n)++;
                        if (p < w.rlog_cap) w.rlog[p] = bid;
                    }
                }
                cand_remove(w.cand_d, w.cand_id, cand_n, bi);
            }
            if (lane == 0) w.sc[SC_STOP] = stop;
        }
        __syncthreads();
        if (w.sc[SC_STOP]) break;
        long long t0 = 0, t1 = 0, t2 = 0;
        if (w.prof && threadIdx.x == 0) t0 = clock64();
        gather_unvisited<DT>(k, g, w, (uint32_t)w.sc[SC_CUR], level, &vis);
        __syncthreads();
        if (w.prof && threadIdx.x == 0) t1 = clock64();
        const int n = w.sc[SC_NBN];
        eval_dists<P, true>(k, w.pivot, n, w.nb_dist, [&](int j, uint32_t &a, uint32_t &) { a = w.nb_ids[j]; });
        __syncthreads();
        if (w.prof && threadIdx.x == 0) {
            t2 = clock64();
            w.prof[0] += t1 - t0;
            w.prof[1] += t2 - t1;
            w.prof[3] = t2; // admit starts
        }
        if (warp0) {
            if (lane == 0) evals += n;
            bool failed = false;
            for (int j0 = 0; j0 < n && !failed; j0 += 32) {
                const int j = j0 + lane;
                const DT dj = j < n ? w.nb_dist[j] : DT(0);
                // whoever fails the test now fails it later too (the bound only shrinks once the set is full)
                unsigned mask = __ballot_sync(0xffffffffu, j < n && (lower > dj || top_n < ef));
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const DT d = __shfl_sync(0xffffffffu, dj, b);
                    if (!(lower > d || top_n < ef)) continue;
                    const uint32_t id = w.nb_ids[j0 + b];
                    if (!cand_insert(w, cand_n, d, id, top_n >= ef, lower)) {
                        failed = true;
                        break;
                    }
                    if (!w.nb_del[j0 + b]) {
                        bool take = true;
                        if (multi) { // emplace by label: a known label keeps its better score, a better one replaces it
                            const uint64_t lab = labels[id];
                            if (regtop) {
                                const int p = reg_find_label(lab);
                                if (p >= 0) {
                                    if (reg_dist_at(p) > d) reg_remove(p);
                                    else take = false;
                                }
                            } else {
                                int p = -1;
                                for (int i0 = 0; i0 < top_n && p < 0; i0 += 32) {
                                    const int i = i0 + lane;
                                    const unsigned m = __ballot_sync(0xffffffffu, i < top_n && labels[w.top_id[i]] == lab);
                                    if (m) p = i0 + __ffs(m) - 1;
                                }
                                if (p >= 0) {
                                    if (w.top_d[p] > d) {
                                        for (int lo = p + 1; lo < top_n; lo += 32) { // shift the tail down by one
                                            const int i = lo + lane;
                                            DT td = DT(0);
                                            uint32_t ti = 0;
                                            if (i < top_n) {
                                                td = w.top_d[i];
                                                ti = w.top_id[i];
                                            }
                                            __syncwarp();
                                            if (i < top_n) {
                                                w.top_d[i - 1] = td;
                                                w.top_id[i - 1] = ti;
                                            }
                                            __syncwarp();
                                        }
                                        top_n--;
                                    } else {
                                        take = false;
                                    }
                                }
                            }
                        }
                        if (take) {
                            if (regtop) reg_insert(d, id);
                            else sorted_insert(w.top_d, w.top_id, top_n, d, id, tl);
                        }
                        if (take && lane == 0 && adm_n < w.adm_cap) {
                            w.adm_d[adm_n] = d;
                            w.adm_id[adm_n] = id;
                        }
                        if (take) adm_n++;
                    }
                    if (top_n > ef) top_n--;
                    if (top_n > 0) lower = regtop ? reg_last() : w.top_d[top_n - 1];
                }
            }
            if (failed) {
                if (lane == 0) w.sc[SC_STATUS] = 1;
                cand_n = 0; // abandon: the caller reruns with a spill area
            }
            if (w.prof && lane == 0) w.prof[2] += clock64() - w.prof[3];
        }
        // nb_* are rewritten only after the next __syncthreads (top of the loop)
    }
    if (regtop && warp0) {
        if (lane < top_n) {
            w.top_d[lane] = t0d;
            w.top_id[lane] = t0i;
        }
        if (lane + 32 < top_n) {
            w.top_d[lane + 32] = t1d;
            w.top_id[lane + 32] = t1i;
        }
    }
    if (threadIdx.x == 0) {
        w.sc[SC_TOPN] = top_n;
        w.sc[SC_ADMN] = adm_n;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
struct SearchArgs {
    KCtx k;
    GraphDev g;
    const uint8_t *q;
    size_t q_stride;
    const float *q_norms;
    const uint64_t *labels;
    uint32_t *visited; // [nq][vis_words]
    size_t vis_words;
    int ef, k_out, cand_cap, max_links;
    void *spill; // per query: spill_cap x (DT, u32) or nullptr
    int spill_cap;
    uint32_t *out_ids;
    void *out_scores;
    uint64_t *out_labels;
    uint32_t *out_counts;
    size_t out_ld;
    uint32_t *status;             // per query
    unsigned long long *counters; // [0] distance evaluations, [1] hops
    size_t pivot_bytes;
    // range flavour
    double radius, epsilon;
    unsigned long long *range_counts; // per query
    size_t range_cap;
    int profile; // VSGPU_HNSW_PROFILE: per-phase cycle counters
    int no_regtop; // VSGPU_HNSW_NO_REGTOP
    int multi;     // HNSWIndex_Multi: result set keyed by label
};

template <typename DT> __device__ __forceinline__ Work<DT> carve(unsigned char *smem, size_t pivot_bytes, int max_links,
                                                                int top_cap, int cand_cap, int adm_cap) {
    Work<DT> w{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char *p = smem + off;
        off += (bytes + 15) / 16 * 16;
        return p;
    };
    w.pivot = take(pivot_bytes);
    w.sc = (int *)take(SC_COUNT * sizeof(int));
    w.sdt = (DT *)take(4 * sizeof(DT));
    w.nb_dist = (DT *)take((size_t)max_links * sizeof(DT));
    w.nb_ids = (uint32_t *)take((size_t)max_links * 4);
    w.nb_del = (uint8_t *)take((size_t)max_links);
    w.top_d = (DT *)take((size_t)top_cap * sizeof(DT));
    w.top_id = (uint32_t *)take((size_t)top_cap * 4);
    w.cand_d = (DT *)take((size_t)cand_cap * sizeof(DT));
    w.cand_id = (uint32_t *)take((size_t)cand_cap * 4);
    w.cand_cap = cand_cap;
    w.adm_d = (DT *)take((size_t)adm_cap * sizeof(DT));
    w.adm_id = (uint32_t *)take((size_t)adm_cap * 4);
    w.adm_cap = adm_cap;
    return w;
}
__host__ __device__ static size_t carve_bytes(size_t dt, size_t pivot_bytes, int max_links, int top_cap, int cand_cap, int adm_cap) {
    auto al = [](size_t b) { return (b + 15) / 16 * 16; };
    return al(pivot_bytes) + al(SC_COUNT * sizeof(int)) + al(4 * dt) + al((size_t)max_links * dt) + al((size_t)max_links * 4) +
           al((size_t)max_links) + al((size_t)top_cap * dt) + al((size_t)top_cap * 4) + al((size_t)cand_cap * dt) + al((size_t)cand_cap * 4) +
           al((size_t)adm_cap * dt) + al((size_t)adm_cap * 4);
}

// entry point + greedy descent to level 1 (searchBottomLayerEP, hnsw.h:1967-1981). Returns false
// for an empty graph. Leaves cur in SC_CUR.
template <class P>
__device__ bool descend(const KCtx &k, const GraphDev &g, Work<typename P::DT> &w, unsigned long long &evals) {
    const int ep = g.state[0], maxl = g.state[1];
    if (ep < 0) return false;
    if (threadIdx.x == 0) {
        w.sc[SC_CUR] = ep;
        w.nb_ids[0] = (uint32_t)ep;
    }
    __syncthreads();
    eval_dists<P, true>(k, w.pivot, 1, w.nb_dist, [&](int j, uint32_t &a, uint32_t &) { a = w.nb_ids[j]; });
    __syncthreads();
    if (threadIdx.x == 0) {
        w.sdt[0] = w.nb_dist[0];
        evals += 1;
    }
    __syncthreads();
    for (int level = maxl; level > 0; level--) greedy_level<P>(k, g, w, level, false, evals);
    return true;
}

template <class P> __global__ void __launch_bounds__(HNSW_THREADS) hnsw_search_kernel(SearchArgs a) {
    using DT = typename P::DT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Work<DT> w = carve<DT>(smem_raw, a.pivot_bytes, a.max_links, a.ef + 1, a.cand_cap, 0);
    const size_t q = blockIdx.x;
    if (a.spill) {
        unsigned char *sp = (unsigned char *)a.spill + q * (size_t)a.spill_cap * (sizeof(DT) + 4);
        w.spill_d = (DT *)sp;
        w.spill_id = (uint32_t *)(sp + (size_t)a.spill_cap * sizeof(DT));
        w.spill_cap = a.spill_cap;
    }
    P::load_pivot(a.k, w.pivot, a.q + q * a.q_stride, a.q_norms ? a.q_norms[q] : 0.f);
    __shared__ long long s_prof[8]; // cycles: 0 gather 1 eval 2 admit 3 scratch 4 bottom layer 5 descent
    if (threadIdx.x < 8) s_prof[threadIdx.x] = 0;
    w.prof = a.profile ? s_prof : nullptr;
    w.no_regtop = a.no_regtop != 0;
    __syncthreads();
    unsigned long long evals = 0, hops = 0;
    int count = 0;
    long long tp = a.profile ? clock64() : 0;
    if (descend<P>(a.k, a.g, w, evals)) {
        if (a.profile && threadIdx.x == 0) {
            s_prof[5] = clock64() - tp;
            tp = clock64();
        }
        Visited vis{a.visited + q * a.vis_words, 0};
        search_layer<P>(a.k, a.g, w, 0, a.ef, a.labels, vis, evals, hops, a.multi != 0);
        if (a.profile && threadIdx.x == 0) s_prof[4] = clock64() - tp;
        count = min(w.sc[SC_TOPN], a.k_out); // ascending (score, label): the k best are the first k
        for (int i = threadIdx.x; i < count; i += blockDim.x) {
            const size_t o = q * a.out_ld + i;
            const uint32_t id = w.top_id[i];
            if (a.out_ids) a.out_ids[o] = id;
            if (a.out_scores) ((DT *)a.out_scores)[o] = w.top_d[i];
            if (a.out_labels) a.out_labels[o] = a.labels[id];
        }
        if (threadIdx.x == 0 && a.status) a.status[q] = (uint32_t)w.sc[SC_STATUS];
    } else if (threadIdx.x == 0 && a.status) {
        a.status[q] = 0;
    }
    for (int j = count + threadIdx.x; j < a.k_out; j += blockDim.x) {
        const size_t o = q * a.out_ld + j;
        if (a.out_ids) a.out_ids[o] = INV;
        if (a.out_scores) ((DT *)a.out_scores)[o] = dt_nan<DT>();
        if (a.out_labels) a.out_labels[o] = ~0ull;
    }
    if (threadIdx.x == 0) {
        if (a.out_counts) a.out_counts[q] = (uint32_t)count;
        if (a.counters) {
            atomicAdd(&a.counters[0], evals);
            atomicAdd(&a.counters[1], hops);
            if (a.profile) {
                for (int i = 0; i < 6; i++) atomicAdd(&a.counters[2 + i], (unsigned long long)s_prof[i]);
                atomicMax(&a.counters[8], hops);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-per-query traversal (top-k, ef <= 64). The CTA-per-query kernel above pays three block-wide barriers per hop and
// leaves seven of its eight warps idle while warp 0 pops and admits: a hop cost 6-11 us for ~2 dependent memory round
// trips of work (r1 profile: barrier stall 21 cycles per issue, issue slots 9 % busy). Here ONE warp owns a query from the
// entry point to the reply — pop, link gather, visited test, distances (eight rows per slice, every load in flight),
// admission — with no barrier at all after the queries are staged, and a CTA carries several independent queries.
// Same algorithm, same order of admissions, same arithmetic: the results are identical (hnsw.h:530-613, 682-721,
// 1210-1258, 1983-2084); anything this kernel does not take (ef > 64, candidate set overflow) goes to the kernel above.
constexpr int WQ_MAX_WARPS = 8;

template <typename DT> struct WarpTop { // the result set in registers: lane L holds entries L and L + 32, ascending
    DT d0 = DT(0), d1 = DT(0);
    uint32_t i0 = 0, i1 = 0;
    int n = 0;
    __device__ __forceinline__ void insert(DT d, uint32_t id, int lane, const TopLess<DT> &tl) {
        const bool p0 = lane < n && tl(d0, i0, d, id);
        const bool p1 = lane + 32 < n && tl(d1, i1, d, id);
        const int cnt = __popc(__ballot_sync(0xffffffffu, p0)) + __popc(__ballot_sync(0xffffffffu, p1));
        const DT u0d = __shfl_up_sync(0xffffffffu, d0, 1);
        const uint32_t u0i = __shfl_up_sync(0xffffffffu, i0, 1);
        DT u1d = __shfl_up_sync(0xffffffffu, d1, 1);
        uint32_t u1i = __shfl_up_sync(0xffffffffu, i1, 1);
        const DT l31d = __shfl_sync(0xffffffffu, d0, 31);
        const uint32_t l31i = __shfl_sync(0xffffffffu, i0, 31);
        if (lane == 0) {
            u1d = l31d;
            u1i = l31i;
        }
        if (lane == cnt) {
            d0 = d;
            i0 = id;
        } else if (lane > cnt) {
            d0 = u0d;
            i0 = u0i;
        }
        if (lane + 32 == cnt) {
            d1 = d;
            i1 = id;
        } else if (lane + 32 > cnt) {
            d1 = u1d;
            i1 = u1i;
        }
        n = min(n + 1, 64);
    }
    __device__ __forceinline__ DT dist_at(int p) const {
        const DT a = __shfl_sync(0xffffffffu, d0, p & 31), b = __shfl_sync(0xffffffffu, d1, p & 31);
        return p < 32 ? a : b;
    }
    __device__ __forceinline__ int find_label(uint64_t lab, int lane, const uint64_t *labels) const {
        const bool p0 = lane < n && labels[i0] == lab;
        const bool p1 = lane + 32 < n && labels[i1] == lab;
        const unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
        return b0 ? __ffs(b0) - 1 : (b1 ? 32 + __ffs(b1) - 1 : -1);
    }
    __device__ __forceinline__ void remove(int p, int lane) {
        const DT e0 = __shfl_down_sync(0xffffffffu, d0, 1), e1 = __shfl_down_sync(0xffffffffu, d1, 1);
        const uint32_t j0 = __shfl_down_sync(0xffffffffu, i0, 1), j1 = __shfl_down_sync(0xffffffffu, i1, 1);
        const DT f1d = __shfl_sync(0xffffffffu, d1, 0);
        const uint32_t f1i = __shfl_sync(0xffffffffu, i1, 0);
        if (lane >= p) {
            d0 = lane == 31 ? f1d : e0;
            i0 = lane == 31 ? f1i : j0;
        }
        if (lane + 32 >= p) {
            d1 = e1;
            i1 = j1;
        }
        n--;
    }
};

struct WarpScratch { // per-warp slices of shared memory
    void *pivot;
    uint32_t *nb_ids;
    uint8_t *nb_del;
    void *nb_dist;
    void *cand_d;
    uint32_t *cand_id;
};
__host__ __device__ inline size_t wq_warp_bytes(size_t dt, size_t pivot_bytes, int max_links, int cand_cap) {
    auto al = [](size_t b) { return (b + 15) / 16 * 16; };
    return al(pivot_bytes) + al((size_t)max_links * 4) + al((size_t)max_links) + al((size_t)max_links * dt) + al((size_t)cand_cap * dt) +
           al((size_t)cand_cap * 4);
}
__device__ __forceinline__ WarpScratch wq_carve(unsigned char *p, size_t dt, size_t pivot_bytes, int max_links, int cand_cap) {
    auto al = [](size_t b) { return (b + 15) / 16 * 16; };
    WarpScratch w;
    w.pivot = p;
    p += al(pivot_bytes);
    w.nb_ids = (uint32_t *)p;
    p += al((size_t)max_links * 4);
    w.nb_del = p;
    p += al((size_t)max_links);
    w.nb_dist = p;
    p += al((size_t)max_links * dt);
    w.cand_d = p;
    p += al((size_t)cand_cap * dt);
    w.cand_id = (uint32_t *)p;
    return w;
}

// links of `node` at `level` that were not visited yet, in link order (warp-wide; see gather_unvisited)
template <typename DT>
__device__ __forceinline__ int wq_gather(const KCtx &k, const GraphDev &g, const WarpScratch &w, uint32_t node, int level,
                                         const Visited *vis, int lane) {
    const uint32_t *rec = links_of(g, node, level);
    const int width = level == 0 ? g.M0 : g.M;
    int base = 0, cnt = 0;
    for (int i0 = 0; i0 < width; i0 += 32) {
        const int i = i0 + lane;
        const uint32_t id = i < width ? rec[1 + i] : INV;
        if (i0 == 0) cnt = (int)rec[0];
        if (i0 >= cnt) break;
        bool take = false;
        uint8_t del = 0;
        if (i < cnt) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(k.rows + (size_t)id * k.row_stride));
            del = g.flags[id];
            take = vis ? !test_and_set(*vis, id) : true;
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (take) {
            const int pos = base + __popc(m & ((1u << lane) - 1));
            w.nb_ids[pos] = id;
            w.nb_del[pos] = del & 1;
        }
        base += __popc(m);
    }
    __syncwarp();
    return base;
}

// nb_dist[j] = dist(row nb_ids[j], query) for j < n (warp-wide)
template <class P, int WIDE> __device__ __forceinline__ void wq_eval(const KCtx &k, const WarpScratch &w, int n, int lane) {
    using DT = typename P::DT;
    DT *out = (DT *)w.nb_dist;
    constexpr int GPW = 32 / P::G;
    const int c = lane % P::G, grp = lane / P::G;
    int base = 0;
    if constexpr (P::HAS_FAST && P::G == 32) {
        if (k.plan.kind == CK_LANES && k.plan.prefix == 0) {
            // eight rows per slice, all their loads in flight before the first FMA. (Tried and dropped, r2: issuing slice
            // i + 1's loads before slice i's FMAs — no gain, the FMAs are too short to hide a memory round trip; and staging a
            // whole hop's rows in shared memory with cp.async before the visited test — 1.46 ms against 0.88 ms per batch of
            // 256: the extra registers and 16 KB per warp cost more occupancy than the overlap returns.)
            auto slice = [&]<int NR>() {
                uint32_t a[NR];
#pragma unroll
                for (int r = 0; r < NR; r++) a[r] = base + r < n ? w.nb_ids[base + r] : INV;
                DT o[NR];
                P::template dists_fast<NR>(k, w.pivot, a, c, o);
                if (lane == 0) {
#pragma unroll
                    for (int r = 0; r < NR; r++)
                        if (base + r < n) out[base + r] = o[r];
                }
                base += NR;
            };
            // WIDE (small batches, latency-bound): sixteen rows per slice while more than eight are left — 64 loads in flight per
            // lane, a hop's ~30 rows cost two memory round trips instead of four (0.72 against 0.81 ms per batch of 256 on the
            // 1 M-node graph) at 204 registers; large batches are throughput-bound and keep the 8-row slices at 128 registers
            // (2.1 M against 1.4 M QPS at batch 4096)
            if constexpr (WIDE) {
                while (n - base > 8) slice.template operator()<16>();
                if (n - base > 4) slice.template operator()<8>();
            } else {
                while (n - base > 4) slice.template operator()<8>();
            }
        }
    }
    constexpr int RU = P::RU;
    for (; base < n; base += GPW * RU) {
        const int j0 = base + grp * RU;
        uint32_t a[RU], b[RU];
#pragma unroll
        for (int r = 0; r < RU; r++) {
            a[r] = j0 + r < n ? w.nb_ids[j0 + r] : INV;
            b[r] = INV;
        }
        DT o[RU];
        P::template dists<true>(k, w.pivot, a, b, c, o);
        if (c == 0) {
#pragma unroll
            for (int r = 0; r < RU; r++)
                if (a[r] != INV) out[j0 + r] = o[r];
        }
    }
    __syncwarp();
}

template <class P, int WIDE> __global__ void __launch_bounds__(WQ_MAX_WARPS * 32) hnsw_search_warp_kernel(SearchArgs a, int wpc, size_t nq) {
    using DT = typename P::DT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t per_warp = wq_warp_bytes(sizeof(DT), a.pivot_bytes, a.max_links, a.cand_cap);
    // stage the CTA's queries: P::load_pivot is CTA-cooperative, one call per warp slot
    for (int ws = 0; ws < wpc; ws++) {
        const size_t qq = (size_t)blockIdx.x * wpc + ws;
        if (qq < nq) P::load_pivot(a.k, smem_raw + (size_t)ws * per_warp, a.q + qq * a.q_stride, a.q_norms ? a.q_norms[qq] : 0.f);
    }
    __syncthreads(); // the only block-wide barrier
    const size_t q = (size_t)blockIdx.x * wpc + warp;
    if (q >= nq) return;
    const WarpScratch w = wq_carve(smem_raw + (size_t)warp * per_warp, sizeof(DT), a.pivot_bytes, a.max_links, a.cand_cap);
    DT *nb_dist = (DT *)w.nb_dist, *cand_d = (DT *)w.cand_d;
    const KCtx &k = a.k;
    const GraphDev &g = a.g;
    unsigned long long evals = 0, hops = 0;
    int count = 0;
    uint32_t status = 0;
    long long prof[6] = {0, 0, 0, 0, 0, 0}; // VSGPU_HNSW_PROFILE cycles: gather, eval, admit, pop, bottom layer, descent
    const long long tstart = a.profile ? clock64() : 0;
    const int ep0 = g.state[0], maxl = g.state[1];
    if (ep0 >= 0) {
        // ---- greedy descent to level 1 (searchBottomLayerEP, greedySearchLevel) ----
        uint32_t cur = (uint32_t)ep0;
        if (lane == 0) w.nb_ids[0] = cur;
        __syncwarp();
        wq_eval<P, WIDE>(k, w, 1, lane);
        DT cur_d = nb_dist[0];
        evals += 1;
        for (int level = maxl; level > 0; level--) {
            for (;;) {
                __syncwarp();
                const int n = wq_gather<DT>(k, g, w, cur, level, nullptr, lane);
                wq_eval<P, WIDE>(k, w, n, lane);
                evals += n;
                bool changed = false;
                for (int j = 0; j < n; j++) { // sequential scan in link order, uniform across the warp
                    const DT dj = nb_dist[j];
                    if (dj < cur_d) {
                        cur_d = dj;
                        cur = w.nb_ids[j];
                        changed = true;
                    }
                }
                if (!changed) break;
            }
        }
        // ---- bottom layer (searchBottomLayer_WithTimeout) ----
        const Visited vis{a.visited + q * a.vis_words, 0};
        const TopLess<DT> tl{a.labels};
        const CandLess<DT> cl;
        const int ef = a.ef;
        const bool multi = a.multi != 0;
        WarpTop<DT> top;
        int cand_n = 0;
        DT lower;
        __syncwarp();
        if (!is_deleted(g, cur)) {
            lower = cur_d;
            top.insert(lower, cur, lane, tl);
        } else {
            lower = dt_max<DT>();
        }
        if (lane == 0) {
            cand_d[0] = lower;
            w.cand_id[0] = cur;
            test_and_set(vis, cur);
        }
        cand_n = 1;
        __syncwarp();
        long long tpop = a.profile ? clock64() : 0;
        const long long tbottom = tpop;
        prof[5] = tpop - tstart;
        for (;;) {
            // pop the best candidate: maximum under pair(-dist, id)
            DT bd = DT(0);
            uint32_t bid = 0;
            int bi = -1;
            for (int i = lane; i < cand_n; i += 32) {
                const DT d = cand_d[i];
                const uint32_t id = w.cand_id[i];
                if (bi < 0 || cl(bd, bid, d, id)) {
                    bd = d;
                    bid = id;
                    bi = i;
                }
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const DT od = __shfl_xor_sync(0xffffffffu, bd, m);
                const uint32_t oid = __shfl_xor_sync(0xffffffffu, bid, m);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
                if (oi >= 0 && (bi < 0 || cl(bd, bid, od, oid))) {
                    bd = od;
                    bid = oid;
                    bi = oi;
                }
            }
            if (bi < 0) break;
            if (bd > lower && top.n >= ef) break;
            __syncwarp();
            if (lane == 0) {
                cand_d[bi] = cand_d[cand_n - 1];
                w.cand_id[bi] = w.cand_id[cand_n - 1];
            }
            cand_n--;
            hops++;
            __syncwarp();
            long long t0 = 0, t1 = 0, t2 = 0;
            if (a.profile) t0 = clock64();
            const int n = wq_gather<DT>(k, g, w, bid, 0, &vis, lane);
            if (a.profile) t1 = clock64();
            wq_eval<P, WIDE>(k, w, n, lane);
            if (a.profile) {
                t2 = clock64();
                prof[0] += t1 - t0;
                prof[1] += t2 - t1;
                prof[3] += t0 - tpop;
            }
            evals += n;
            bool failed = false;
            for (int j0 = 0; j0 < n && !failed; j0 += 32) {
                const int j = j0 + lane;
                const DT dj = j < n ? nb_dist[j] : DT(0);
                // whoever fails the test now fails it later too (the bound only shrinks once the set is full)
                unsigned mask = __ballot_sync(0xffffffffu, j < n && (lower > dj || top.n < ef));
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const DT d = __shfl_sync(0xffffffffu, dj, b);
                    if (!(lower > d || top.n < ef)) continue;
                    const uint32_t id = w.nb_ids[j0 + b];
                    if (cand_n >= a.cand_cap) {
                        if (top.n >= ef) { // drop entries that can never be expanded (see cand_insert)
                            int out = 0;
                            for (int base = 0; base < cand_n; base += 32) {
                                const int i = base + lane;
                                DT td = DT(0);
                                uint32_t ti = 0;
                                bool keep = false;
                                if (i < cand_n) {
                                    td = cand_d[i];
                                    ti = w.cand_id[i];
                                    keep = !(td > lower);
                                }
                                const unsigned m = __ballot_sync(0xffffffffu, keep);
                                __syncwarp();
                                if (keep) {
                                    const int pos = out + __popc(m & ((1u << lane) - 1));
                                    cand_d[pos] = td;
                                    w.cand_id[pos] = ti;
                                }
                                __syncwarp();
                                out += __popc(m);
                            }
                            cand_n = out;
                        }
                        if (cand_n >= a.cand_cap) {
                            failed = true;
                            break;
                        }
                    }
                    if (lane == 0) {
                        cand_d[cand_n] = d;
                        w.cand_id[cand_n] = id;
                    }
                    __syncwarp();
                    cand_n++;
                    if (!w.nb_del[j0 + b]) {
                        bool take = true;
                        if (multi) {
                            const int p = top.find_label(a.labels[id], lane, a.labels);
                            if (p >= 0) {
                                if (top.dist_at(p) > d) top.remove(p, lane);
                                else take = false;
                            }
                        }
                        if (take) top.insert(d, id, lane, tl);
                    }
                    if (top.n > ef) top.n--;
                    if (top.n > 0) lower = top.dist_at(top.n - 1);
                }
            }
            if (failed) {
                status = 1; // the host redoes this query on the CTA kernel with a spill area
                break;
            }
            if (a.profile) {
                const long long t3 = clock64();
                prof[2] += t3 - t2;
                tpop = t3;
            }
        }
        if (a.profile) prof[4] = clock64() - tbottom;
        count = min(top.n, a.k_out);
        const size_t o0 = q * a.out_ld;
        if (lane < count) {
            if (a.out_ids) a.out_ids[o0 + lane] = top.i0;
            if (a.out_scores) ((DT *)a.out_scores)[o0 + lane] = top.d0;
            if (a.out_labels) a.out_labels[o0 + lane] = a.labels[top.i0];
        }
        if (lane + 32 < count) {
            if (a.out_ids) a.out_ids[o0 + lane + 32] = top.i1;
            if (a.out_scores) ((DT *)a.out_scores)[o0 + lane + 32] = top.d1;
            if (a.out_labels) a.out_labels[o0 + lane + 32] = a.labels[top.i1];
        }
    }
    for (int j = count + lane; j < a.k_out; j += 32) {
        const size_t o = q * a.out_ld + j;
        if (a.out_ids) a.out_ids[o] = INV;
        if (a.out_scores) ((DT *)a.out_scores)[o] = dt_nan<DT>();
        if (a.out_labels) a.out_labels[o] = ~0ull;
    }
    if (lane == 0) {
        if (a.status) a.status[q] = status;
        if (a.out_counts) a.out_counts[q] = (uint32_t)count;
        if (a.counters) {
            atomicAdd(&a.counters[0], evals);
            atomicAdd(&a.counters[1], hops);
            if (a.profile) {
                for (int i = 0; i < 6; i++) atomicAdd(&a.counters[2 + i], (unsigned long long)prof[i]);
                atomicMax(&a.counters[8], hops);
            }
        }
    }
}

// Range search at level 0 (hnsw.h:2086-2150 + :615-680). Results are appended unordered to
// out_*[q * range_cap ...]; range_counts[q] is the number found (may exceed the capacity: the host
// retries with a larger buffer — the traversal is deterministic).
template <class P> __global__ void __launch_bounds__(HNSW_THREADS) hnsw_range_kernel(SearchArgs a) {
    using DT = typename P::DT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Work<DT> w = carve<DT>(smem_raw, a.pivot_bytes, a.max_links, 1, a.cand_cap, 0);
    const size_t q = blockIdx.x;
    if (a.spill) {
        unsigned char *sp = (unsigned char *)a.spill + q * (size_t)a.spill_cap * (sizeof(DT) + 4);
        w.spill_d = (DT *)sp;
        w.spill_id = (uint32_t *)(sp + (size_t)a.spill_cap * sizeof(DT));
        w.spill_cap = a.spill_cap;
    }
    P::load_pivot(a.k, w.pivot, a.q + q * a.q_stride, a.q_norms ? a.q_norms[q] : 0.f);
    __syncthreads();
    unsigned long long evals = 0, hops = 0, found = 0;
    if (!descend<P>(a.k, a.g, w, evals)) {
        if (threadIdx.x == 0) a.range_counts[q] = 0;
        return;
    }
    const DT radius = (DT)a.radius;
    Visited vis{a.visited + q * a.vis_words, 0};
    CandLess<DT> cl;
    const bool warp0 = threadIdx.x < 32;
    const int lane = threadIdx.x & 31;
    int cand_n = 0; // warp 0's (uniform) copies
    DT dyn = DT(0), bound = DT(0);
    auto emit = [&](uint32_t id, DT d) { // lane 0
        if (found < a.range_cap) {
            const size_t o = q * a.range_cap + found;
            if (a.out_ids) a.out_ids[o] = id;
            if (a.out_scores) ((DT *)a.out_scores)[o] = d;
            if (a.out_labels) a.out_labels[o] = a.labels[id];
        }
        found++;
    };
    if (warp0) {
        const uint32_t ep = (uint32_t)w.sc[SC_CUR];
        DT ep_dist;
        if (lane == 0) w.sc[SC_STATUS] = 0;
        if (is_deleted(a.g, ep)) {
            ep_dist = dt_max<DT>();
            bound = dyn = ep_dist;
        } else {
            ep_dist = w.sdt[0]; // distance of the greedy descent's final node
            dyn = ep_dist;
            if (ep_dist <= radius) {
                if (lane == 0) emit(ep, ep_dist);
                dyn = radius;
            }
            bound = (DT)((double)dyn * (1.0 + a.epsilon));
        }
        cand_insert(w, cand_n, ep_dist, ep, false, bound);
        if (lane == 0) test_and_set(vis, ep);
    }
    for (;;) {
        if (warp0) {
            int stop = 0;
            DT cd = DT(0);
            uint32_t bid = 0;
            const int bi = cand_n ? cand_best(w.cand_d, w.cand_id, cand_n, cd, bid) : -1;
            if (bi < 0 || cd > bound) stop = 1;
            else {
                if (lane == 0) {
                    w.sc[SC_CUR] = (int)bid;
                    hops++;
                }
                cand_remove(w.cand_d, w.cand_id, cand_n, bi);
                if (cd < dyn && cd >= radius) {
                    dyn = cd;
                    bound = (DT)((double)dyn * (1.0 + a.epsilon));
                }
            }
            if (lane == 0) w.sc[SC_STOP] = stop;
        }
        __syncthreads();
        if (w.sc[SC_STOP]) break;
        gather_unvisited<DT>(a.k, a.g, w, (uint32_t)w.sc[SC_CUR], 0, &vis);
        __syncthreads();
        const int n = w.sc[SC_NBN];
        eval_dists<P, true>(a.k, w.pivot, n, w.nb_dist, [&](int j, uint32_t &x, uint32_t &) { x = w.nb_ids[j]; });
        __syncthreads();
        if (warp0) {
            if (lane == 0) evals += n;
            bool failed = false;
            for (int j0 = 0; j0 < n && !failed; j0 += 32) {
                const int j = j0 + lane;
                const DT dj = j < n ? w.nb_dist[j] : DT(0);
                unsigned mask = __ballot_sync(0xffffffffu, j < n && dj < bound); // the bound is fixed during a hop
                while (mask) {
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const DT d = __shfl_sync(0xffffffffu, dj, b);
                    const uint32_t id = w.nb_ids[j0 + b];
                    if (!cand_insert(w, cand_n, d, id, true, bound)) {
                        failed = true;
                        break;
                    }
                    if (lane == 0 && d <= radius && !w.nb_del[j0 + b]) emit(id, d);
                }
            }
            if (failed) {
                if (lane == 0) w.sc[SC_STATUS] = 1;
                cand_n = 0;
            }
        }
    }
    if (threadIdx.x == 0) {
        a.range_counts[q] = found;
        if (a.status) a.status[q] = (uint32_t)w.sc[SC_STATUS];
        if (a.counters) {
            atomicAdd(&a.counters[0], evals);
            atomicAdd(&a.counters[1], hops);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Batch iterator (hnsw_batch_iterator.h:93-267, hnsw_single_batch_iterator.h:36-78): one resumable
// traversal whose state — visited bitmap, the candidates min-heap on (dist, id), the "extras"
// min-heap on (dist, label), lower bound, entry point — lives in HBM between calls. Every call
// returns the next n results exactly as HNSW_BatchIterator::getNextResults does.
struct BiState {
    int entry;          // bottom-layer entry point, -1 = not computed yet, -2 = empty graph
    int depleted;
    unsigned long long returned;
    unsigned long long cand_n, extra_n;
    double lower;       // lower_bound (DistType widened)
};
struct BiArgs {
    KCtx k;
    GraphDev g;
    const uint8_t *q;
    float q_norm;
    const uint64_t *labels;
    uint32_t *visited;
    BiState *state;
    void *cand_d;       // [cap] DistType
    uint32_t *cand_id;
    void *extra_d;
    uint32_t *extra_id;
    void *top_d;        // [ef + 1]
    uint32_t *top_id;
    int ef, n_res, max_links;
    unsigned long long label_count;
    uint32_t *out_ids;
    void *out_scores;
    uint64_t *out_labels;
    uint32_t *out_count;
    size_t pivot_bytes;
    unsigned long long *counters;
    // HNSWMulti_BatchIterator (hnsw_multi_batch_iterator.h:39-99): the result set is keyed by label (a flat array here:
    // emplace = find the label, keep the better score) and rows whose label was handed out by an earlier batch are
    // skipped — the host marks every row of a returned label in `returned_bits` between calls
    int multi;
    const uint32_t *returned_bits;
};
// std::greater on pair<DistType, key>: min-heap. Expressed as the "less" our max-heap helpers take.
template <typename DT> struct MinLess {
    const uint64_t *labels; // nullptr: key = id
    __device__ __forceinline__ bool operator()(DT ad, uint32_t ai, DT bd, uint32_t bi) const {
        // a "less" than b in heap order  <=>  pair(a) > pair(b)
        if (bd < ad) return true;
        if (ad < bd) return false;
        if (labels) return labels[bi] < labels[ai];
        return bi < ai;
    }
};

template <class P> __global__ void __launch_bounds__(HNSW_THREADS) hnsw_bi_kernel(BiArgs a) {
    using DT = typename P::DT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Work<DT> w = carve<DT>(smem_raw, a.pivot_bytes, a.max_links, 1, 1, 0);
    P::load_pivot(a.k, w.pivot, a.q, a.q_norm);
    __syncthreads();
    unsigned long long evals = 0, hops = 0;
    BiState *st = a.state;
    // first call: greedy descent to the bottom-layer entry point
    if (st->entry == -1) {
        const bool ok = descend<P>(a.k, a.g, w, evals);
        __syncthreads();
        if (threadIdx.x == 0) st->entry = ok ? w.sc[SC_CUR] : -2;
        __syncthreads();
    }
    DT *cand_d = (DT *)a.cand_d, *extra_d = (DT *)a.extra_d, *top_d = (DT *)a.top_d;
    TopLess<DT> tl{a.labels};
    MinLess<DT> cl{nullptr}, el{a.labels};
    int top_n = 0, cand_n = 0, extra_n = 0;
    DT lower = DT(0);
    const int ef = a.ef;
    bool skip_scan = false;
    // multi-value flavour, thread 0 only: `top` as a flat array (one entry per label)
    auto was_returned = [&](uint32_t id) { return (a.returned_bits[id >> 5] >> (id & 31)) & 1u; };
    auto m_max = [&]() { // index of the largest entry under (score, label)
        int m = 0;
        for (int i = 1; i < top_n; i++)
            if (tl(top_d[m], a.top_id[m], top_d[i], a.top_id[i])) m = i;
        return m;
    };
    auto m_remove = [&](int i) {
        top_d[i] = top_d[top_n - 1];
        a.top_id[i] = a.top_id[top_n - 1];
        top_n--;
    };
    auto m_emplace = [&](DT d, uint32_t id) {
        const uint64_t lab = a.labels[id];
        for (int i = 0; i < top_n; i++)
            if (a.labels[a.top_id[i]] == lab) {
                if (top_d[i] > d) {
                    top_d[i] = d;
                    a.top_id[i] = id;
                }
                return;
            }
        top_d[top_n] = d;
        a.top_id[top_n] = id;
        top_n++;
    };
    if (threadIdx.x == 0) {
        cand_n = (int)st->cand_n;
        extra_n = (int)st->extra_n;
        lower = (DT)st->lower;
        int stop = 0;
        if (st->entry < 0) {
            st->depleted = 1;
            stop = 2; // nothing to scan, nothing to return
        } else {
            if (st->returned == 0 && extra_n == 0 && cand_n == 0) {
                w.sc[SC_AUX0] = 1; // need the entry point's distance
            } else {
                w.sc[SC_AUX0] = 0;
            }
        }
        w.sc[SC_STOP] = stop;
        w.nb_ids[0] = (uint32_t)max(st->entry, 0);
    }
    __syncthreads();
    if (w.sc[SC_STOP] == 2) {
        if (threadIdx.x == 0) *a.out_count = 0;
        return;
    }
    if (w.sc[SC_AUX0]) {
        eval_dists<P, true>(a.k, w.pivot, 1, w.nb_dist, [&](int j, uint32_t &x, uint32_t &) { x = w.nb_ids[j]; });
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t ep = w.nb_ids[0];
            lower = is_deleted(a.g, ep) ? dt_max<DT>() : w.nb_dist[0];
            a.visited[ep >> 5] |= 1u << (ep & 31);
            heap_push(cand_d, a.cand_id, cand_n, lower, ep, cl);
            evals += 1;
        }
    }
    if (threadIdx.x == 0) {
        // fillFromExtras
        while (top_n < ef && extra_n > 0) {
            if (a.multi) {
                if (!was_returned(a.extra_id[0])) m_emplace(extra_d[0], a.extra_id[0]);
            } else {
                heap_push(top_d, a.top_id, top_n, extra_d[0], a.extra_id[0], tl);
            }
            heap_pop(extra_d, a.extra_id, extra_n, el);
        }
        skip_scan = top_n == ef;
        w.sc[SC_AUX1] = skip_scan ? 1 : 0;
    }
    __syncthreads();
    if (!w.sc[SC_AUX1]) {
        Visited vis{a.visited, 0};
        for (;;) {
            if (threadIdx.x == 0) {
                int stop = 0;
                if (cand_n == 0) stop = 1;
                else if (cand_d[0] > lower && top_n >= ef) stop = 1;
                else {
                    const DT d = cand_d[0];
                    const uint32_t id = a.cand_id[0];
                    if (!is_deleted(a.g, id) && a.multi) {
                        // updateHeaps, label-keyed (hnsw_multi_batch_iterator.h:59-84)
                        if ((lower > d || top_n < ef) && !was_returned(id)) {
                            m_emplace(d, id);
                            if (top_n > ef) {
                                const int m = m_max();
                                heap_push(extra_d, a.extra_id, extra_n, top_d[m], a.top_id[m], el);
                                m_remove(m);
                            }
                            lower = top_d[m_max()];
                        }
                    } else if (!is_deleted(a.g, id)) {
                        // updateHeaps
                        if (top_n < ef) {
                            heap_push(top_d, a.top_id, top_n, d, id, tl);
                            lower = top_d[0];
                        } else if (lower > d) {
                            heap_push(top_d, a.top_id, top_n, d, id, tl);
                            heap_push(extra_d, a.extra_id, extra_n, top_d[0], a.top_id[0], el);
                            heap_pop(top_d, a.top_id, top_n, tl);
                            lower = top_d[0];
                        }
                    }
                    heap_pop(cand_d, a.cand_id, cand_n, cl);
                    w.sc[SC_CUR] = (int)id;
                    hops++;
                }
                w.sc[SC_STOP] = stop;
            }
            __syncthreads();
            if (w.sc[SC_STOP]) break;
            gather_unvisited<DT>(a.k, a.g, w, (uint32_t)w.sc[SC_CUR], 0, &vis);
            __syncthreads();
            const int n = w.sc[SC_NBN];
            eval_dists<P, true>(a.k, w.pivot, n, w.nb_dist, [&](int j, uint32_t &x, uint32_t &) { x = w.nb_ids[j]; });
            __syncthreads();
            if (threadIdx.x == 0) {
                evals += n;
                for (int j = 0; j < n; j++) heap_push(cand_d, a.cand_id, cand_n, w.nb_dist[j], w.nb_ids[j], cl);
            }
        }
    }
    if (threadIdx.x == 0) {
        if (!skip_scan && top_n < ef) st->depleted = 1;
        // prepareResults: spare results go back to the extras
        while (top_n > a.n_res) {
            const int m = a.multi ? m_max() : 0;
            heap_push(extra_d, a.extra_id, extra_n, top_d[m], a.top_id[m], el);
            if (a.multi) m_remove(m);
            else heap_pop(top_d, a.top_id, top_n, tl);
        }
        const int count = top_n;
        for (int i = count - 1; i >= 0; i--) {
            const int m = a.multi ? m_max() : 0;
            const uint32_t id = a.top_id[m];
            a.out_ids[i] = id;
            ((DT *)a.out_scores)[i] = top_d[m];
            a.out_labels[i] = a.labels[id];
            if (a.multi) m_remove(m);
            else heap_pop(top_d, a.top_id, top_n, tl);
        }
        *a.out_count = (uint32_t)count;
        st->returned += (unsigned long long)count;
        if (st->returned == a.label_count) st->depleted = 1;
        st->cand_n = (unsigned long long)cand_n;
        st->extra_n = (unsigned long long)extra_n;
        st->lower = (double)lower;
        if (a.counters) {
            atomicAdd(&a.counters[0], evals);
            atomicAdd(&a.counters[1], hops);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Builder
struct InsertArgs {
    KCtx k;
    GraphDev g;
    uint32_t first, n; // elements [first, first + n) are already in the store
    uint32_t *tags;    // [capacity] visited tags
    uint32_t *tag_counter;
    int efc, cand_cap, max_links;
    void *spill;
    int spill_cap;
    size_t pivot_bytes;
    unsigned long long *counters;
    uint32_t *status;
    int rv_warps; // warps that run revisits side by side (0: the CTA-wide path)
    // Batched builder. mode 0: sequential (one CTA inserts [first, first + n) one after the other). mode 1: CTA b
    // searches for element first + b on the graph as it stands and records, per level, the neighbours it selected and
    // every node whose links it read. mode 2: one CTA commits the recorded elements in order while their read sets are
    // untouched by the earlier commits of the round (the traversal would then have been step-for-step identical), and
    // stops at the first one that is not: the host re-runs mode 1 from there.
    int mode;
    uint32_t *bitmaps;    // [slots][bm_words] visited sets (mode 1)
    size_t bm_words;
    uint32_t *rlog;       // [slots][rlog_cap]
    int rlog_cap;
    int *res_meta;        // [slots][4 + BATCH_MAX_LEVELS]: ok, entry snapshot, max-level snapshot, rlog count, ns per level
    uint32_t *res_id;     // [slots][BATCH_MAX_LEVELS][M]
    void *res_d;          // [slots][BATCH_MAX_LEVELS][M] DistType
    uint32_t *mod_stamp;  // [capacity] id of the last inserted element that rewrote the node's links
    uint32_t *committed;  // mode 2: number of elements committed
};
constexpr int BATCH_MAX_LEVELS = 16;

// Scratch of the neighbour-selection heuristic, carved after the Work arrays
template <typename DT> struct Heur {
    DT *sd;        // sorted candidate distances [cap]
    uint32_t *sid; // ids
    uint32_t *spos; // original position of each sorted candidate
    uint8_t *keep; // per sorted candidate: selected
    uint32_t *sel; // selected sorted positions [maxM]
    DT *pd;        // [HEUR_CHUNK][maxM + HEUR_CHUNK]
    DT *in_d;      // unsorted input [cap]
    uint32_t *in_id;
    int cap, maxM;
};
__host__ __device__ static size_t heur_bytes(size_t dt, int cap, int maxM) {
    auto al = [](size_t b) { return (b + 15) / 16 * 16; };
    return al(cap * dt) + 2 * al((size_t)cap * 4) + al(cap) + al((size_t)maxM * 4) +
           al((size_t)HEUR_CHUNK * (maxM + HEUR_CHUNK) * dt) + al(cap * dt) + al((size_t)cap * 4);
}
template <typename DT> __device__ __forceinline__ Heur<DT> carve_heur(unsigned char *p, int cap, int maxM) {
    Heur<DT> h{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char *r = p + off;
        off += (bytes + 15) / 16 * 16;
        return r;
    };
    h.sd = (DT *)take(cap * sizeof(DT));
    h.sid = (uint32_t *)take((size_t)cap * 4);
    h.spos = (uint32_t *)take((size_t)cap * 4);
    h.keep = (uint8_t *)take(cap);
    h.sel = (uint32_t *)take((size_t)maxM * 4);
    h.pd = (DT *)take((size_t)HEUR_CHUNK * (maxM + HEUR_CHUNK) * sizeof(DT));
    h.in_d = (DT *)take(cap * sizeof(DT));
    h.in_id = (uint32_t *)take((size_t)cap * 4);
    h.cap = cap;
    h.maxM = maxM;
    return h;
}

// sort h.in_* [n] ascending by (dist, id) into h.sd/sid/spos (rank sort, whole CTA)
template <typename DT> __device__ void rank_sort(const Heur<DT> &h, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const DT d = h.in_d[i];
        const uint32_t id = h.in_id[i];
        int rank = 0;
        for (int j = 0; j < n; j++) {
            const DT dj = h.in_d[j];
            rank += (dj < d || (dj == d && h.in_id[j] < id)) ? 1 : 0;
        }
        h.sd[rank] = d;
        h.sid[rank] = id;
        h.spos[rank] = (uint32_t)i;
    }
    __syncthreads();
}

// getNeighborsByHeuristic2 (hnsw.h:741-799) over the sorted candidates: a candidate is kept unless an
// already kept one is closer to it than the query is. Candidates are examined HEUR_CHUNK at a time:
// the distances of a chunk to everything it could be compared with are evaluated together
// (speculatively, they have no side effects) and one thread then applies the sequential rule.
// Result: h.keep[] per sorted position, h.sel[0..ns) in order; returns ns through sc[SC_AUX2].
template <class P> __device__ void heuristic(const KCtx &k, const Heur<typename P::DT> &h, int *sc, int n, int maxM) {
    using DT = typename P::DT;
    for (int i = threadIdx.x; i < n; i += blockDim.x) h.keep[i] = 0;
    if (threadIdx.x == 0) sc[SC_AUX2] = 0;
    __syncthreads();
    for (int pos = 0; pos < n; pos += HEUR_CHUNK) {
        const int ns0 = sc[SC_AUX2];
        if (ns0 >= maxM) break;
        const int C = min(HEUR_CHUNK, n - pos);
        const int cols = ns0 + C;
        for (int j = threadIdx.x; j < C * cols; j += blockDim.x) {
            const int ca = j / cols, col = j % cols;
            uint32_t a = INV, b = INV;
            if (col < ns0) {
                a = h.sid[pos + ca];
                b = h.sid[h.sel[col]];
            } else if (col - ns0 < ca) {
                a = h.sid[pos + ca];
                b = h.sid[pos + col - ns0];
            }
            if (a != INV) h.pd[j] = P::dist_thread(k, a, b);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int ns = ns0;
            for (int ca = 0; ca < C && ns < maxM; ca++) {
                const DT dq = h.sd[pos + ca];
                bool good = true;
                for (int i = 0; i < ns && good; i++) {
                    const int col = i < ns0 ? i : ns0 + ((int)h.sel[i] - pos);
                    if (h.pd[ca * cols + col] < dq) good = false;
                }
                if (good) {
                    h.keep[pos + ca] = 1;
                    h.sel[ns++] = (uint32_t)(pos + ca);
                }
            }
            sc[SC_AUX2] = ns;
        }
        __syncthreads();
    }
}

// Per-warp scratch of revisitNeighborConnections: the (<= M0 + 1) candidates, their sorted order and
// the lower triangle of their pairwise distances.
template <typename DT> struct RvScratch {
    DT *cd, *sd, *D;
    uint32_t *cid, *sid, *spos, *kept;
    uint8_t *keep;
};
__host__ __device__ static size_t rv_bytes(size_t dt, int nc) {
    auto al = [](size_t b) { return (b + 15) / 16 * 16; };
    return 2 * al((size_t)nc * dt) + al((size_t)nc * nc * dt) + 4 * al((size_t)nc * 4) + al((size_t)nc);
}
template <typename DT> __device__ __forceinline__ RvScratch<DT> carve_rv(unsigned char *p, int nc) {
    RvScratch<DT> r{};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        unsigned char *q = p + off;
        off += (bytes + 15) / 16 * 16;
        return q;
    };
    r.cd = (DT *)take((size_t)nc * sizeof(DT));
    r.sd = (DT *)take((size_t)nc * sizeof(DT));
    r.D = (DT *)take((size_t)nc * nc * sizeof(DT));
    r.cid = (uint32_t *)take((size_t)nc * 4);
    r.sid = (uint32_t *)take((size_t)nc * 4);
    r.spos = (uint32_t *)take((size_t)nc * 4);
    r.kept = (uint32_t *)take((size_t)nc * 4);
    r.keep = (uint8_t *)take((size_t)nc);
    return r;
}

// revisitNeighborConnections (hnsw.h:801-868) for ONE full neighbour, by ONE warp: the neighbour's links
// plus the new element compete for its max_M slots under getNeighborsByHeuristic2. The revisits of an
// insertion touch disjoint link lists, so the warps of the CTA run them side by side; inside, every
// lane evaluates whole (row, row) distances (dist_thread): all candidate-to-neighbour distances, then
// the lower triangle of candidate-to-candidate distances, then the sequential keep/drop rule is a
// ballot per candidate.
template <class P>
__device__ bool warp_revisit(const KCtx &k, const GraphDev &g, const RvScratch<typename P::DT> &r, uint32_t e, uint32_t nb,
                             typename P::DT d_nb, volatile uint32_t *nb_rec, int maxM) {
    using DT = typename P::DT;
    const int lane = threadIdx.x & 31;
    const int cnt = (int)nb_rec[0];
    const int nc = cnt + 1;
    for (int j = lane; j < cnt; j += 32) {
        const uint32_t id = nb_rec[1 + j];
        r.cid[1 + j] = id;
        r.cd[1 + j] = P::dist_thread(k, id, nb);
    }
    if (lane == 0) {
        r.cid[0] = e;
        r.cd[0] = d_nb;
    }
    __syncwarp();
    for (int i = lane; i < nc; i += 32) {
        const DT d = r.cd[i];
        const uint32_t id = r.cid[i];
        int rank = 0;
        for (int j = 0; j < nc; j++) {
            const DT dj = r.cd[j];
            rank += (dj < d || (dj == d && r.cid[j] < id)) ? 1 : 0;
        }
        r.sd[rank] = d;
        r.sid[rank] = id;
        r.spos[rank] = (uint32_t)i;
        r.keep[i] = 0;
    }
    __syncwarp();
    const int npairs = nc * (nc - 1) / 2;
    for (int p = lane; p < npairs; p += 32) {
        int i = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
        while (i * (i - 1) / 2 > p) i--;
        while ((i + 1) * i / 2 <= p) i++;
        const int j = p - i * (i - 1) / 2;
        r.D[i * nc + j] = P::dist_thread(k, r.sid[i], r.sid[j]);
    }
    __syncwarp();
    int nkeep = 0;
    for (int i = 0; i < nc && nkeep < maxM; i++) {
        const DT dq = r.sd[i];
        bool bad = false;
        for (int k0 = 0; k0 < nkeep; k0 += 32) {
            const int kk = k0 + lane;
            const bool b = kk < nkeep && r.D[i * nc + r.kept[kk]] < dq;
            if (__any_sync(0xffffffffu, b)) {
                bad = true;
                break;
            }
        }
        if (!bad) {
            if (lane == 0) {
                r.kept[nkeep] = (uint32_t)i;
                r.keep[r.spos[i]] = 1;
            }
            nkeep++;
            __syncwarp();
        }
    }
    __syncwarp();
    int changed = 0;
    if (lane == 0) {
        // keep flags by original position: 0 = the new element, 1 + j = link j
        int out = 0;
        for (int j = 0; j < cnt; j++)
            if (r.keep[1 + j]) nb_rec[1 + out++] = nb_rec[1 + j];
        changed = out != cnt; // an old link was dropped ...
        if (r.keep[0] && out < maxM) {
            nb_rec[1 + out++] = e; // (the caller appends nb to the new element's list)
            changed = 1;          // ... or the new element got in
        }
        nb_rec[0] = (uint32_t)out;
    }
    __syncwarp();
    (void)g;
    return __shfl_sync(0xffffffffu, changed, 0) != 0;
}

template <class P> __global__ void __launch_bounds__(HNSW_THREADS, 1) hnsw_insert_kernel(InsertArgs a) {
    using DT = typename P::DT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = a.g.M, M0 = a.g.M0;
    const int top_cap = a.efc + 1;
    Work<DT> w = carve<DT>(smem_raw, a.pivot_bytes, a.max_links, top_cap, a.cand_cap, M);
    const size_t work_bytes = carve_bytes(sizeof(DT), a.pivot_bytes, a.max_links, top_cap, a.cand_cap, M);
    const int hcap = max(top_cap, M0 + 1);
    Heur<DT> h = carve_heur<DT>(smem_raw + work_bytes, hcap, M0);
    const size_t rv_off = work_bytes + heur_bytes(sizeof(DT), hcap, M0);
    const size_t rv_each = rv_bytes(sizeof(DT), M0 + 1);
    DT *const cand_d0 = w.cand_d;
    uint32_t *const cand_id0 = w.cand_id;
    const int cand_cap0 = w.cand_cap;
    if (a.spill && a.mode != 1) { // one spill area: not for the concurrent search CTAs
        w.spill_d = (DT *)a.spill;
        w.spill_id = (uint32_t *)((unsigned char *)a.spill + (size_t)a.spill_cap * sizeof(DT));
        w.spill_cap = a.spill_cap;
    }
    unsigned long long evals = 0, hops = 0;
    __shared__ uint32_t s_tag;
    __shared__ long long s_prof[8]; // 0 gather 1 eval 2 admit 3 scratch 4 search_layer total 5 select 6 connect 7 greedy+setup
    if (threadIdx.x < 8) s_prof[threadIdx.x] = 0;
    w.prof = s_prof;
    long long tp = 0;

    __shared__ int s_rlog_n, s_valid;
    const bool per_cta = a.mode == 1;
    const uint32_t e_begin = per_cta ? a.first + blockIdx.x : a.first;
    const uint32_t e_end = per_cta ? e_begin + 1 : a.first + a.n;
    for (uint32_t e = e_begin; e < e_end; e++) {
        __syncthreads();
        const uint32_t slot = e - a.first;
        int *meta = a.res_meta ? a.res_meta + (size_t)slot * (4 + BATCH_MAX_LEVELS) : nullptr;
        const bool commit_mode = a.mode == 2;
        // mode 3 works from the state its search saw (other CTAs may be raising the entry point right now)
        const int ep = a.g.state[0], maxl = a.g.state[1];
        const int lvl = (int)a.g.levels[e];
        if (a.mode == 1) {
            if (threadIdx.x == 0) {
                s_rlog_n = 0;
                meta[0] = (ep >= 0 && min(lvl, maxl) < BATCH_MAX_LEVELS) ? 1 : 0;
                meta[1] = ep;
                meta[2] = maxl;
            }
            w.rlog = a.rlog + (size_t)slot * a.rlog_cap;
            w.rlog_cap = a.rlog_cap;
            w.rlog_n = &s_rlog_n;
            __syncthreads();
            if (!meta[0]) return;
        }
        if (a.mode == 2) {
            // valid while the graph state and every link list this element's search read are as it saw them
            if (threadIdx.x == 0) s_valid = (meta[0] == 1 && meta[1] == ep && meta[2] == maxl && meta[3] <= a.rlog_cap) ? 1 : 0;
            __syncthreads();
            if (s_valid) {
                const uint32_t *log = a.rlog + (size_t)slot * a.rlog_cap;
                const int ln = meta[3];
                int bad = 0;
                for (int i = threadIdx.x; i < ln; i += blockDim.x) bad |= (a.mod_stamp[log[i]] >= a.first) ? 1 : 0;
                if (bad) s_valid = 0; // benign race: all writers store 0
            }
            __syncthreads();
            if (!s_valid) {
                if (threadIdx.x == 0) *a.committed = slot;
                return;
            }
        }
        if (ep < 0) { // first element: nothing to connect to
            __syncthreads();
            if (threadIdx.x == 0) {
                a.g.state[0] = (int)e;
                a.g.state[1] = lvl;
            }
            continue;
        }
        if (!commit_mode) P::load_pivot(a.k, w.pivot, a.k.rows + (size_t)e * a.k.row_stride, a.k.norms ? a.k.norms[e] : 0.f);
        if (threadIdx.x == 0) {
            w.sc[SC_CUR] = ep;
            w.nb_ids[0] = (uint32_t)ep;
        }
        __syncthreads();
        int max_common = maxl;
        if (lvl < maxl) max_common = lvl;
        if (lvl < maxl && !commit_mode) {
            eval_dists<P, true>(a.k, w.pivot, 1, w.nb_dist, [&](int j, uint32_t &x, uint32_t &) { x = w.nb_ids[j]; });
            __syncthreads();
            if (threadIdx.x == 0) {
                w.sdt[0] = w.nb_dist[0];
                evals += 1;
            }
            __syncthreads();
            for (int level = maxl; level > lvl; level--) greedy_level<P>(a.k, a.g, w, level, true, evals);
        }
        for (int level = max_common; level >= 0; level--) {
            const int maxMcur = level ? M : M0;
            DT *sel_d = w.top_d; // selected neighbours (distance, id) live in the result arrays once the search is over
            uint32_t *sel_id = w.top_id;
            int ns = 0;
            if (commit_mode) {
                // load what the search CTA selected for this level
                ns = meta[4 + level];
                if (threadIdx.x == 0) tp = clock64();
                for (int i = threadIdx.x; i < ns; i += blockDim.x) {
                    sel_id[i] = a.res_id[((size_t)slot * BATCH_MAX_LEVELS + level) * M + i];
                    sel_d[i] = ((const DT *)a.res_d)[((size_t)slot * BATCH_MAX_LEVELS + level) * M + i];
                }
                __syncthreads();
                if (ns == 0) continue;
            } else {
            // fresh visited set; candidate set back in shared memory
            if (a.mode == 1) {
                uint32_t *bm = a.bitmaps + (size_t)slot * a.bm_words;
                for (size_t i = threadIdx.x; i < a.bm_words; i += blockDim.x) bm[i] = 0;
            } else if (threadIdx.x == 0) {
                uint32_t t = *a.tag_counter + 1;
                if (t == 0) t = 1; // the host clears the tag array before the counter can wrap
                *a.tag_counter = t;
                s_tag = t;
            }
            w.cand_d = cand_d0;
            w.cand_id = cand_id0;
            w.cand_cap = cand_cap0;
            __syncthreads();
            Visited vis{a.mode == 1 ? a.bitmaps + (size_t)slot * a.bm_words : a.tags, a.mode == 1 ? 0u : s_tag};
            if (threadIdx.x == 0) tp = clock64();
            search_layer<P>(a.k, a.g, w, level, a.efc, nullptr, vis, evals, hops);
            if (threadIdx.x == 0) {
                s_prof[4] += clock64() - tp;
                tp = clock64();
            }
            if (w.sc[SC_STATUS]) {
                // candidate set overflow without a spill area: sequential mode reports it, a search CTA just
                // invalidates its slot (the host falls back to the sequential path for that element)
                if (threadIdx.x == 0) {
                    if (a.mode == 1) meta[0] = 0;
                    else *a.status = 1;
                }
                return;
            }
            const int n = w.sc[SC_TOPN];
            if (n == 0) { // entry point was marked deleted and nothing else was reachable
                if (a.mode == 1 && threadIdx.x == 0) meta[4 + level] = 0;
                continue;
            }

            // ---- choose the new element's neighbours (mutuallyConnectNewElement :870-890) ----
            if (n < M) {
                // fewer than M candidates: all are kept, in the order of the result heap's underlying
                // array = the admissions replayed through push_heap (no pop ever happened)
                if (threadIdx.x == 0) {
                    TopLess<DT> tl{nullptr};
                    int m = 0;
                    for (int i = 0; i < n; i++) heap_push(h.sd, h.sid, m, w.adm_d[i], w.adm_id[i], tl);
                    int best = 0;
                    for (int i = 1; i < n; i++)
                        if (h.sd[i] < h.sd[best]) best = i;
                    w.sc[SC_AUX3] = (int)h.sid[best]; // next closest entry point
                    for (int i = 0; i < n; i++) h.sel[i] = (uint32_t)i;
                    w.sc[SC_AUX2] = n;
                }
                __syncthreads();
                ns = n;
            } else {
                // the result set is already sorted ascending by (dist, id)
                for (int i = threadIdx.x; i < n; i += blockDim.x) {
                    h.sd[i] = w.top_d[i];
                    h.sid[i] = w.top_id[i];
                    h.spos[i] = (uint32_t)i;
                }
                __syncthreads();
                heuristic<P>(a.k, h, w.sc, n, M);
                ns = w.sc[SC_AUX2];
                if (threadIdx.x == 0) w.sc[SC_AUX3] = (int)h.sid[h.sel[0]];
                __syncthreads();
            }
            // the selected list (distance, id), copied out of the heuristic scratch (revisit reuses it)
            __syncthreads();
            if (threadIdx.x == 0) {
                // gather first (sel positions ascend, so in-place reads stay ahead of writes only via a copy)
                for (int i = 0; i < ns; i++) {
                    w.nb_dist[i] = h.sd[h.sel[i]];
                    w.nb_ids[i] = h.sid[h.sel[i]];
                }
                for (int i = 0; i < ns; i++) {
                    sel_d[i] = w.nb_dist[i];
                    sel_id[i] = w.nb_ids[i];
                }
            }
            __syncthreads();
            if (a.mode == 1) {
                // record the selection; the next level starts from the closest selected neighbour
                for (int i = threadIdx.x; i < ns; i += blockDim.x) {
                    a.res_id[((size_t)slot * BATCH_MAX_LEVELS + level) * M + i] = sel_id[i];
                    ((DT *)a.res_d)[((size_t)slot * BATCH_MAX_LEVELS + level) * M + i] = sel_d[i];
                }
                if (threadIdx.x == 0) {
                    meta[4 + level] = ns;
                    w.sc[SC_CUR] = w.sc[SC_AUX3];
                }
                __syncthreads();
                continue;
            }
            } // !commit_mode
            uint32_t *new_rec = links_of(a.g, e, level);
            if (threadIdx.x == 0) {
                s_prof[5] += clock64() - tp;
                tp = clock64();
            }
            if (a.rv_warps > 0) {
                // The selected neighbours own disjoint link lists and the new element always has room for all of
                // them (ns <= M <= max_M), so the per-neighbour updates are independent: one warp each.
                const int warp = threadIdx.x >> 5;
                if (warp < a.rv_warps) {
                    const RvScratch<DT> rv = carve_rv<DT>(smem_raw + rv_off + (size_t)warp * rv_each, M0 + 1);
                    for (int si = warp; si < ns; si += a.rv_warps) {
                        const uint32_t nb = sel_id[si];
                        if (is_deleted(a.g, nb)) continue;
                        volatile uint32_t *nb_rec = links_of(a.g, nb, level);
                        bool changed = true;
                        if ((int)nb_rec[0] < maxMcur) {
                            if ((threadIdx.x & 31) == 0) {
                                nb_rec[1 + nb_rec[0]] = e;
                                nb_rec[0] = nb_rec[0] + 1;
                            }
                            __syncwarp();
                        } else {
                            changed = warp_revisit<P>(a.k, a.g, rv, e, nb, sel_d[si], nb_rec, maxMcur);
                        }
                        // a full neighbour that rejects the new element and keeps all its links is untouched: later
                        // elements of the round that read its list are still valid
                        if (a.mod_stamp && changed && (threadIdx.x & 31) == 0) a.mod_stamp[nb] = e;
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    uint32_t c = 0;
                    for (int si = 0; si < ns && (int)c < maxMcur; si++)
                        if (!is_deleted(a.g, sel_id[si])) new_rec[1 + c++] = sel_id[si];
                    new_rec[0] = c;
                }
            } else {
                for (int si = 0; si < ns; si++) {
                    __syncthreads();
                    const uint32_t nb = sel_id[si];
                    uint32_t *nb_rec = links_of(a.g, nb, level);
                    if (threadIdx.x == 0) {
                        int action = 0; // 0 nothing/simple, 1 revisit, 2 stop
                        if ((int)new_rec[0] == maxMcur) action = 2;
                        else if (is_deleted(a.g, nb)) action = 0;
                        else if ((int)nb_rec[0] < maxMcur) {
                            new_rec[1 + new_rec[0]] = nb;
                            new_rec[0]++;
                            nb_rec[1 + nb_rec[0]] = e;
                            nb_rec[0]++;
                        } else action = 1;
                        if (a.mod_stamp && action != 2 && !is_deleted(a.g, nb)) a.mod_stamp[nb] = e;
                        w.sc[SC_AUX1] = action;
                    }
                    __syncthreads();
                    const int action = w.sc[SC_AUX1];
                    if (action == 2) break;
                    if (action == 0) continue;
                    // ---- revisitNeighborConnections (:801-868) ----
                    const int cnt = (int)nb_rec[0];
                    const int nc = cnt + 1;
                    eval_dists<P, false>(a.k, nullptr, cnt, h.in_d + 1, [&](int j, uint32_t &x, uint32_t &y) {
                        x = nb_rec[1 + j];
                        y = nb;
                    });
                    for (int j = threadIdx.x; j < cnt; j += blockDim.x) h.in_id[1 + j] = nb_rec[1 + j];
                    if (threadIdx.x == 0) {
                        h.in_d[0] = sel_d[si];
                        h.in_id[0] = e;
                        evals += cnt;
                    }
                    __syncthreads();
                    rank_sort<DT>(h, nc);
                    heuristic<P>(a.k, h, w.sc, nc, maxMcur);
                    if (threadIdx.x == 0) {
                        // keep flags by original position: 0 = the new element, 1 + j = link j
                        bool new_chosen = false;
                        int kept = 0;
                        // sorted position -> original position; mark in nb_ids scratch
                        for (int i = 0; i < nc; i++) w.nb_ids[h.spos[i]] = h.keep[i];
                        new_chosen = w.nb_ids[0] != 0;
                        for (int j = 0; j < cnt; j++)
                            if (w.nb_ids[1 + j]) nb_rec[1 + kept++] = nb_rec[1 + j];
                        if ((int)new_rec[0] < maxMcur && !is_deleted(a.g, nb)) {
                            new_rec[1 + new_rec[0]] = nb;
                            new_rec[0]++;
                            if (new_chosen && kept < maxMcur) nb_rec[1 + kept++] = e;
                        }
                        nb_rec[0] = (uint32_t)kept;
                    }
                    __syncthreads();
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) s_prof[6] += clock64() - tp;
            if (threadIdx.x == 0 && a.mode == 0) w.sc[SC_CUR] = w.sc[SC_AUX3];
            __syncthreads();
        }
        __syncthreads();
        if (a.mode == 1) {
            if (threadIdx.x == 0) meta[3] = s_rlog_n;
            break;
        }
        if (threadIdx.x == 0 && lvl > maxl) {
            a.g.state[0] = (int)e;
            a.g.state[1] = lvl;
        }
        if (threadIdx.x == 0 && a.mode == 2) *a.committed = slot + 1;
        __threadfence_block();
    }
    if (threadIdx.x == 0 && a.counters) {
        atomicAdd(&a.counters[0], evals);
        atomicAdd(&a.counters[1], hops);
        if (a.mode != 1)
            for (int i = 0; i < 8; i++) atomicAdd(&a.counters[2 + i], (unsigned long long)s_prof[i]);
    }
}

} // namespace vsgpu

// ================================================================================================
// host side of the C-ABI
using namespace vsgpu;

struct vsgpu_hnsw {
    vsgpu_store *s = nullptr;
    int M = 0, M0 = 0, efc = 0;
    bool multi = false;    // several rows per label: queries key their result set by label
    size_t capacity = 0;   // nodes the arrays are sized for
    size_t count = 0;      // nodes in the graph
    uint32_t *l0 = nullptr, *up = nullptr, *up_off = nullptr, *levels = nullptr, *tags = nullptr, *tag_counter = nullptr;
    uint8_t *flags = nullptr;
    int *state = nullptr;
    size_t up_records = 0, up_capacity = 0;
    int entry = -1, max_level = -1;
    Scratch visited, spill, out, misc;
    unsigned long long *counters = nullptr; // [0] evals [1] hops
    uint32_t *status = nullptr;
    unsigned long long last_evals = 0, last_hops = 0;
    unsigned long long last_prof[8] = {0};
    // batched builder state
    uint32_t *mod_stamp = nullptr;
    size_t mod_cap = 0;
    Scratch batch; // bitmaps, read logs, per-slot results
    uint32_t *committed = nullptr;
    unsigned long long rounds = 0, round_elems = 0;
    float last_ms = 0;
    uint32_t host_tag = 0;
};

namespace vsgpu {

static KCtx make_kctx(const vsgpu_store *s) {
    KCtx k{};
    k.rows = s->rows;
    k.row_stride = s->row_stride;
    k.type = s->type;
    k.metric = s->metric;
    k.plan = s->plan;
    k.norms = s->has_norm ? s->norms : nullptr;
    k.chunks = (int)(s->row_stride / 16);
    return k;
}
static GraphDev make_graph(const vsgpu_hnsw *g) {
    GraphDev d{};
    d.l0 = g->l0;
    d.up = g->up;
    d.up_off = g->up_off;
    d.levels = g->levels;
    d.flags = g->flags;
    d.state = g->state;
    d.M = g->M;
    d.M0 = g->M0;
    return d;
}

template <typename F> static int dispatch_policy(const vsgpu_store *s, F &&f) {
    const ChainPlan &p = s->plan;
    const bool l2 = p.is_l2;
    if (p.kind == CK_INT) return s->type == VSGPU_UINT8 ? f.template operator()<PolInt<true>>() : f.template operator()<PolInt<false>>();
    if (p.kind == CK_SEQ) return s->type == VSGPU_FLOAT64 ? f.template operator()<PolSeq<double>>() : f.template operator()<PolSeq<float>>();
    if (s->type == VSGPU_FLOAT64)
        return l2 ? f.template operator()<PolChain<double, 16, false, true, VSGPU_FLOAT64>>()
                  : f.template operator()<PolChain<double, 16, false, false, VSGPU_FLOAT64>>();
    if (p.kind == CK_BF16_DP) return f.template operator()<PolChain<float, 16, true, false, VSGPU_BFLOAT16>>();
    if (p.kind == CK_BF16_VBMI2) return f.template operator()<PolChain<float, 16, false, true, VSGPU_BFLOAT16>>();
    if (s->type == VSGPU_FLOAT16)
        return l2 ? f.template operator()<PolChain<float, 32, false, true, VSGPU_FLOAT16>>()
                  : f.template operator()<PolChain<float, 32, false, false, VSGPU_FLOAT16>>();
    return l2 ? f.template operator()<PolChain<float, 32, false, true, VSGPU_FLOAT32>>()
              : f.template operator()<PolChain<float, 32, false, false, VSGPU_FLOAT32>>();
}

template <typename T> static int regrow(T *&p, size_t old_n, size_t new_n, cudaStream_t st, bool zero) {
    T *np = nullptr;
    VS_CUDA(cudaMalloc(&np, new_n * sizeof(T)));
    if (zero) VS_CUDA(cudaMemsetAsync(np, 0, new_n * sizeof(T), st));
    if (p && old_n) VS_CUDA(cudaMemcpyAsync(np, p, old_n * sizeof(T), cudaMemcpyDeviceToDevice, st));
    VS_CUDA(cudaStreamSynchronize(st));
    if (p) cudaFree(p);
    p = np;
    return VSGPU_OK;
}

static int graph_reserve(vsgpu_hnsw *g, size_t nodes, size_t up_records) {
    vsgpu_store *s = g->s;
    if (nodes > g->capacity) {
        size_t cap = std::max<size_t>(nodes, g->capacity + g->capacity / 2);
        cap = std::max<size_t>(cap, 1024);
        VS_TRY(regrow(g->l0, g->capacity * (g->M0 + 1), cap * (g->M0 + 1), s->stream, true));
        VS_TRY(regrow(g->up_off, g->capacity, cap, s->stream, true));
        VS_TRY(regrow(g->levels, g->capacity, cap, s->stream, true));
        VS_TRY(regrow(g->tags, g->capacity, cap, s->stream, true));
        VS_TRY(regrow(g->flags, g->capacity, cap, s->stream, true));
        g->capacity = cap;
    }
    if (up_records > g->up_capacity) {
        size_t cap = std::max<size_t>(up_records, g->up_capacity + g->up_capacity / 2);
        cap = std::max<size_t>(cap, 256);
        VS_TRY(regrow(g->up, g->up_capacity * (g->M + 1), cap * (g->M + 1), s->stream, true));
        g->up_capacity = cap;
    }
    return VSGPU_OK;
}

static int read_state(vsgpu_hnsw *g) {
    int st[2];
    VS_CUDA(cudaMemcpyAsync(st, g->state, sizeof(st), cudaMemcpyDeviceToHost, g->s->stream));
    VS_CUDA(cudaStreamSynchronize(g->s->stream));
    g->entry = st[0];
    g->max_level = st[1];
    return VSGPU_OK;
}

static size_t smem_limit(int device) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess || v <= 0) v = 227 * 1024;
    return (size_t)v;
}

} // namespace vsgpu

extern "C" {

vsgpu_hnsw *vsgpu_hnsw_create(vsgpu_store *s, size_t M, size_t ef_construction) {
    if (!s || M < 2 || 2 * M > 512) {
        set_error("vsgpu_hnsw_create: M must be in [2, 256]");
        return nullptr;
    }
    if (cudaSetDevice(s->device) != cudaSuccess) return nullptr;
    auto *g = new vsgpu_hnsw();
    g->s = s;
    g->M = (int)M;
    g->M0 = (int)(2 * M);
    g->efc = (int)std::max(ef_construction, M);
    bool ok = cudaMalloc(&g->state, 2 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaMalloc(&g->counters, 16 * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMalloc(&g->status, sizeof(uint32_t)) == cudaSuccess;
    ok = ok && cudaMalloc(&g->tag_counter, sizeof(uint32_t)) == cudaSuccess;
    if (ok) {
        const int st[2] = {-1, -1};
        ok = cudaMemcpy(g->state, st, sizeof(st), cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMemset(g->tag_counter, 0, sizeof(uint32_t)) == cudaSuccess;
        ok = ok && cudaDeviceSynchronize() == cudaSuccess; // default-stream work: the store's stream does not wait for it
    }
    if (!ok) {
        set_error("vsgpu_hnsw_create: device allocation failed");
        vsgpu_hnsw_destroy(g);
        return nullptr;
    }
    return g;
}

void vsgpu_hnsw_destroy(vsgpu_hnsw *g) {
    if (!g) return;
    cudaSetDevice(g->s->device);
    cudaStreamSynchronize(g->s->stream);
    for (void *p : {(void *)g->l0, (void *)g->up, (void *)g->up_off, (void *)g->levels, (void *)g->tags, (void *)g->flags,
                    (void *)g->state, (void *)g->counters, (void *)g->status, (void *)g->tag_counter})
        if (p) cudaFree(p);
    for (Scratch *sc : {&g->visited, &g->spill, &g->out, &g->misc, &g->batch})
        if (sc->ptr) cudaFree(sc->ptr);
    if (g->mod_stamp) cudaFree(g->mod_stamp);
    if (g->committed) cudaFree(g->committed);
    delete g;
}

size_t vsgpu_hnsw_size(const vsgpu_hnsw *g) { return g->count; }
void vsgpu_hnsw_set_multi(vsgpu_hnsw *g, int multi) { g->multi = multi != 0; }
size_t vsgpu_hnsw_device_bytes(const vsgpu_hnsw *g) {
    return g->capacity * ((size_t)(g->M0 + 1) * 4 + 4 + 4 + 4 + 1) + g->up_capacity * (size_t)(g->M + 1) * 4 +
           g->visited.bytes + g->spill.bytes + g->out.bytes + g->misc.bytes;
}
int vsgpu_hnsw_entry(const vsgpu_hnsw *g, long *entry, long *max_level) {
    if (entry) *entry = g->entry;
    if (max_level) *max_level = g->max_level;
    return VSGPU_OK;
}

int vsgpu_hnsw_insert(vsgpu_hnsw *g, size_t n, const uint32_t *levels) {
    vsgpu_store *s = g->s;
    if (n == 0) return VSGPU_OK;
    if (!levels || g->count + n > s->count) {
        set_error("vsgpu_hnsw_insert: rows must be appended to the store first");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    size_t add_records = 0;
    std::vector<uint32_t> offs(n);
    for (size_t i = 0; i < n; i++) {
        offs[i] = (uint32_t)(g->up_records + add_records);
        add_records += levels[i];
    }
    VS_TRY(graph_reserve(g, std::max(g->count + n, s->capacity), g->up_records + add_records));
    VS_CUDA(cudaMemcpyAsync(g->levels + g->count, levels, n * 4, cudaMemcpyHostToDevice, s->stream));
    VS_CUDA(cudaMemcpyAsync(g->up_off + g->count, offs.data(), n * 4, cudaMemcpyHostToDevice, s->stream));
    // new nodes start with empty link lists (arrays are zeroed when grown, but ids may be reused after import)
    VS_CUDA(cudaMemsetAsync(g->l0 + g->count * (g->M0 + 1), 0, n * (size_t)(g->M0 + 1) * 4, s->stream));
    if (add_records)
        VS_CUDA(cudaMemsetAsync(g->up + g->up_records * (g->M + 1), 0, add_records * (size_t)(g->M + 1) * 4, s->stream));
    VS_CUDA(cudaMemsetAsync(g->flags + g->count, 0, n, s->stream));
    // visited tags: one per search_layer call; clear before a 32-bit wrap could alias
    if ((unsigned long long)g->host_tag + (unsigned long long)n * 64ull > 0xfffffff0ull) {
        VS_CUDA(cudaMemsetAsync(g->tags, 0, g->capacity * 4, s->stream));
        VS_CUDA(cudaMemsetAsync(g->tag_counter, 0, 4, s->stream));
        g->host_tag = 0;
    }
    VS_CUDA(cudaMemsetAsync(g->status, 0, 4, s->stream));
    VS_CUDA(cudaMemsetAsync(g->counters, 0, 128, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream)); // offs goes out of scope
    const size_t dt = s->type == VSGPU_FLOAT64 ? 8 : 4;
    VS_TRY(ensure_scratch(s, g->spill, (g->capacity + 1) * (dt + 4)));
    InsertArgs a{};
    a.k = make_kctx(s);
    a.g = make_graph(g);
    a.tags = g->tags;
    a.tag_counter = g->tag_counter;
    a.efc = g->efc;
    a.cand_cap = 2 * g->efc + 64;
    a.max_links = g->M0 + 1;
    a.spill = g->spill.ptr;
    a.spill_cap = (int)std::min<size_t>(g->capacity, 0x7fffffff);
    a.counters = g->counters;
    a.status = g->status;
    // batched builder buffers (slots = one search CTA per SM)
    // Small graphs: almost every pair of concurrent searches shares a modified hub, rounds commit ~1 element and the
    // extra launches cost more than they save (measured: 20 k rows 1.17 vs 1.02 ms per insert; 100 k rows 0.79 vs 1.01).
    // VSGPU_HNSW_SEQ_BUILD / VSGPU_HNSW_BATCH_BUILD force either path (A/B runs, tests).
    const bool seq_only = getenv("VSGPU_HNSW_SEQ_BUILD") != nullptr ||
                          (g->count + n < 50000 && getenv("VSGPU_HNSW_BATCH_BUILD") == nullptr);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
    const size_t slots = (size_t)std::max(sms, 1);
    const size_t bm_words = (g->capacity + 31) / 32;
    const int rlog_cap = 4096;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_bm = 0, o_log = al(slots * bm_words * 4), o_meta = o_log + al(slots * (size_t)rlog_cap * 4),
                 o_rid = o_meta + al(slots * (4 + BATCH_MAX_LEVELS) * 4), o_rd = o_rid + al(slots * BATCH_MAX_LEVELS * (size_t)g->M * 4),
                 batch_bytes = o_rd + al(slots * BATCH_MAX_LEVELS * (size_t)g->M * dt);
    if (!seq_only) {
        VS_TRY(ensure_scratch(s, g->batch, batch_bytes));
        if (g->mod_cap < g->capacity) {
            VS_TRY(regrow(g->mod_stamp, g->mod_cap, g->capacity, s->stream, true));
            g->mod_cap = g->capacity;
        }
        if (!g->committed) VS_CUDA(cudaMalloc(&g->committed, 4));
        uint8_t *bb = (uint8_t *)g->batch.ptr;
        a.bitmaps = (uint32_t *)(bb + o_bm);
        a.bm_words = bm_words;
        a.rlog = (uint32_t *)(bb + o_log);
        a.rlog_cap = rlog_cap;
        a.res_meta = (int *)(bb + o_meta);
        a.res_id = (uint32_t *)(bb + o_rid);
        a.res_d = bb + o_rd;
        a.mod_stamp = g->mod_stamp;
        a.committed = g->committed;
    }
    uint32_t st = 0;
    const int rc = dispatch_policy(s, [&]<class P>() -> int {
        a.pivot_bytes = P::pivot_bytes(s);
        size_t smem = carve_bytes(dt, a.pivot_bytes, a.max_links, a.efc + 1, a.cand_cap, g->M) +
                      heur_bytes(dt, std::max(a.efc + 1, g->M0 + 1), g->M0);
        const size_t rv_each = rv_bytes(dt, g->M0 + 1);
        a.rv_warps = 0;
        if (smem < smem_limit(s->device)) a.rv_warps = (int)std::min<size_t>(HNSW_THREADS / 32, (smem_limit(s->device) - smem) / rv_each);
        if (const char *ev = getenv("VSGPU_HNSW_RV_WARPS")) a.rv_warps = std::min(a.rv_warps, std::max(0, atoi(ev))); // tests: force the CTA-wide path
        smem += (size_t)a.rv_warps * rv_each;
        if (smem > smem_limit(s->device)) {
            set_error("vsgpu_hnsw_insert: efConstruction / dim too large for the shared-memory builder");
            return (int)VSGPU_ERR_ARG;
        }
        auto kern = hnsw_insert_kernel<P>;
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VS_CUDA(cudaEventRecord(s->ev0, s->stream));
        auto sequential = [&](uint32_t first, uint32_t cnt) -> int {
            a.mode = 0;
            a.first = first;
            a.n = cnt;
            kern<<<1, HNSW_THREADS, smem, s->stream>>>(a);
            VS_CUDA(cudaGetLastError());
            return (int)VSGPU_OK;
        };
        uint32_t done = 0;
        const uint32_t base = (uint32_t)g->count;
        if (seq_only) {
            VS_TRY(sequential(base, (uint32_t)n));
            done = (uint32_t)n;
        }
        while (done < n) {
            const uint32_t first = base + done;
            if (first == 0 || (g->entry < 0 && done == 0)) { // the first element of an empty graph has nothing to search
                VS_TRY(sequential(first, 1));
                done += 1;
                continue;
            }
            const uint32_t bsz = (uint32_t)std::min<size_t>(n - done, slots);
            a.first = first;
            a.n = bsz;
            a.mode = 1; // search: one CTA per element, on the graph as it stands
            kern<<<bsz, HNSW_THREADS, smem, s->stream>>>(a);
            VS_CUDA(cudaGetLastError());
            VS_CUDA(cudaMemsetAsync(g->committed, 0, 4, s->stream));
            a.mode = 2; // commit in order while the read sets are intact
            kern<<<1, HNSW_THREADS, smem, s->stream>>>(a);
            VS_CUDA(cudaGetLastError());
            uint32_t committed = 0;
            VS_CUDA(cudaMemcpyAsync(&committed, g->committed, 4, cudaMemcpyDeviceToHost, s->stream));
            VS_CUDA(cudaMemcpyAsync(&st, g->status, 4, cudaMemcpyDeviceToHost, s->stream));
            VS_CUDA(cudaStreamSynchronize(s->stream));
            if (st) break;
            g->rounds++;
            g->round_elems += committed;
            if (committed == 0) { // the slot could not be recorded (level cap, candidate overflow): insert it sequentially
                VS_TRY(sequential(first, 1));
                committed = 1;
            }
            done += committed;
        }
        VS_CUDA(cudaEventRecord(s->ev1, s->stream));
        return (int)VSGPU_OK;
    });
    VS_TRY(rc);
    unsigned long long ctr[10] = {0};
    VS_CUDA(cudaMemcpyAsync(&st, g->status, 4, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaMemcpyAsync(ctr, g->counters, 80, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    cudaEventElapsedTime(&g->last_ms, s->ev0, s->ev1);
    g->last_evals = ctr[0];
    g->last_hops = ctr[1];
    for (int i = 0; i < 8; i++) g->last_prof[i] = ctr[2 + i];
    if (getenv("VSGPU_HNSW_PROFILE"))
        fprintf(stderr, "[vsgpu_hnsw_insert] n=%zu ms=%.2f evals=%llu hops=%llu rounds=%llu (%.1f committed per round) cycles: gather=%llu eval=%llu admit=%llu search=%llu select=%llu connect=%llu\n", n,
                g->last_ms, ctr[0], ctr[1], g->rounds, g->rounds ? (double)g->round_elems / (double)g->rounds : 0.0, ctr[2], ctr[3], ctr[4], ctr[6], ctr[7], ctr[8]);
    if (st != 0) {
        set_error("vsgpu_hnsw_insert: candidate set overflow");
        return VSGPU_ERR_OVERFLOW;
    }
    g->host_tag += (uint32_t)(n * 64);
    g->count += n;
    g->up_records += add_records;
    return read_state(g);
}

int vsgpu_hnsw_import(vsgpu_hnsw *g, size_t n, const uint32_t *levels, const uint32_t *l0, const uint32_t *upper,
                      size_t upper_records, long entry, long max_level) {
    vsgpu_store *s = g->s;
    if (n > s->count) {
        set_error("vsgpu_hnsw_import: more nodes than rows in the store");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    std::vector<uint32_t> offs(n);
    size_t rec = 0;
    for (size_t i = 0; i < n; i++) {
        offs[i] = (uint32_t)rec;
        rec += levels[i];
    }
    if (rec != upper_records) {
        set_error("vsgpu_hnsw_import: upper_records != sum(levels)");
        return VSGPU_ERR_ARG;
    }
    VS_TRY(graph_reserve(g, std::max(n, s->capacity), rec));
    if (n) {
        VS_CUDA(cudaMemcpyAsync(g->levels, levels, n * 4, cudaMemcpyHostToDevice, s->stream));
        VS_CUDA(cudaMemcpyAsync(g->up_off, offs.data(), n * 4, cudaMemcpyHostToDevice, s->stream));
        VS_CUDA(cudaMemcpyAsync(g->l0, l0, n * (size_t)(g->M0 + 1) * 4, cudaMemcpyHostToDevice, s->stream));
        if (rec) VS_CUDA(cudaMemcpyAsync(g->up, upper, rec * (size_t)(g->M + 1) * 4, cudaMemcpyHostToDevice, s->stream));
        VS_CUDA(cudaMemsetAsync(g->flags, 0, n, s->stream));
    }
    const int st[2] = {(int)entry, (int)max_level};
    VS_CUDA(cudaMemcpyAsync(g->state, st, sizeof(st), cudaMemcpyHostToDevice, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    g->count = n;
    g->up_records = rec;
    g->entry = (int)entry;
    g->max_level = (int)max_level;
    return VSGPU_OK;
}

int vsgpu_hnsw_export(const vsgpu_hnsw *g, uint32_t *levels, uint32_t *l0, uint32_t *upper, size_t upper_cap_records,
                      size_t *upper_records) {
    vsgpu_store *s = g->s;
    VS_CUDA(cudaSetDevice(s->device));
    if (upper_records) *upper_records = g->up_records;
    const size_t n = g->count;
    if (n == 0) return VSGPU_OK;
    if (levels) VS_CUDA(cudaMemcpyAsync(levels, g->levels, n * 4, cudaMemcpyDeviceToHost, s->stream));
    if (l0) VS_CUDA(cudaMemcpyAsync(l0, g->l0, n * (size_t)(g->M0 + 1) * 4, cudaMemcpyDeviceToHost, s->stream));
    if (upper && g->up_records) {
        if (upper_cap_records < g->up_records) {
            set_error("vsgpu_hnsw_export: upper buffer too small");
            return VSGPU_ERR_OVERFLOW;
        }
        VS_CUDA(cudaMemcpyAsync(upper, g->up, g->up_records * (size_t)(g->M + 1) * 4, cudaMemcpyDeviceToHost, s->stream));
    }
    VS_CUDA(cudaStreamSynchronize(s->stream));
    return VSGPU_OK;
}

int vsgpu_hnsw_set_deleted(vsgpu_hnsw *g, size_t id, int deleted) {
    if (id >= g->count) {
        set_error("vsgpu_hnsw_set_deleted: bad id");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(g->s->device));
    const uint8_t v = deleted ? 1 : 0;
    VS_CUDA(cudaMemcpyAsync(g->flags + id, &v, 1, cudaMemcpyHostToDevice, g->s->stream));
    VS_CUDA(cudaStreamSynchronize(g->s->stream));
    return VSGPU_OK;
}

/* One node's link records read back to the host: level 0 record (2M + 1 words: count, links) followed by one
 * (M + 1)-word record per upper level. Debug / introspection path (VecSimDebug_GetElementNeighborsInHNSWGraph). */
int vsgpu_hnsw_node(const vsgpu_hnsw *g, size_t id, uint32_t *level_out, uint32_t *records, size_t cap_words) {
    vsgpu_store *s = g->s;
    if (id >= g->count || !level_out) {
        set_error("vsgpu_hnsw_node: bad id");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    uint32_t lvl = 0, off = 0;
    VS_CUDA(cudaMemcpyAsync(&lvl, g->levels + id, 4, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaMemcpyAsync(&off, g->up_off + id, 4, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    *level_out = lvl;
    const size_t need = (size_t)(g->M0 + 1) + (size_t)lvl * (g->M + 1);
    if (!records) return VSGPU_OK;
    if (cap_words < need) {
        set_error("vsgpu_hnsw_node: buffer too small");
        return VSGPU_ERR_OVERFLOW;
    }
    VS_CUDA(cudaMemcpyAsync(records, g->l0 + id * (size_t)(g->M0 + 1), (size_t)(g->M0 + 1) * 4, cudaMemcpyDeviceToHost, s->stream));
    if (lvl)
        VS_CUDA(cudaMemcpyAsync(records + (g->M0 + 1), g->up + (size_t)off * (g->M + 1), (size_t)lvl * (g->M + 1) * 4,
                                cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    return VSGPU_OK;
}

int vsgpu_hnsw_last_stats(const vsgpu_hnsw *g, unsigned long long *dist_evals, unsigned long long *hops, float *ms) {
    if (dist_evals) *dist_evals = g->last_evals;
    if (hops) *hops = g->last_hops;
    if (ms) *ms = g->last_ms;
    return VSGPU_OK;
}

} // extern "C"

namespace vsgpu {

// queries staged on the device (see stage_queries_device); outputs on the device
static int hnsw_search_core(vsgpu_hnsw *g, const void *q, size_t nq, size_t q_stride, const float *q_norms, size_t k,
                            size_t ef, bool range, double radius, double epsilon, size_t range_cap, uint32_t *out_ids,
                            void *out_scores, uint64_t *out_labels, uint32_t *out_counts, unsigned long long *range_counts,
                            uint32_t *status, bool with_spill) {
    vsgpu_store *s = g->s;
    const size_t dt = s->type == VSGPU_FLOAT64 ? 8 : 4;
    const size_t vis_words = (g->count + 31) / 32;
    VS_TRY(ensure_scratch(s, g->visited, nq * vis_words * 4 + 256));
    VS_CUDA(cudaMemsetAsync(g->visited.ptr, 0, nq * vis_words * 4, s->stream));
    SearchArgs a{};
    a.k = make_kctx(s);
    a.g = make_graph(g);
    a.q = (const uint8_t *)q;
    a.q_stride = q_stride;
    a.q_norms = q_norms;
    a.labels = s->labels;
    a.visited = (uint32_t *)g->visited.ptr;
    a.vis_words = vis_words;
    a.ef = (int)ef;
    a.k_out = (int)k;
    a.max_links = g->M0;
    a.out_ids = out_ids;
    a.out_scores = out_scores;
    a.out_labels = out_labels;
    a.out_counts = out_counts;
    a.out_ld = k;
    a.status = status;
    a.counters = g->counters;
    a.radius = radius;
    a.epsilon = epsilon;
    a.range_counts = range_counts;
    a.range_cap = range_cap;
    a.profile = getenv("VSGPU_HNSW_PROFILE") != nullptr;
    a.no_regtop = getenv("VSGPU_HNSW_NO_REGTOP") != nullptr;
    a.multi = g->multi ? 1 : 0;
    if (with_spill) {
        VS_TRY(ensure_scratch(s, g->spill, nq * (g->count + 1) * (dt + 4)));
        a.spill = g->spill.ptr;
        a.spill_cap = (int)std::min<size_t>(g->count + 1, 0x7fffffff);
    }
    return dispatch_policy(s, [&]<class P>() -> int {
        a.pivot_bytes = P::pivot_bytes(s);
        const size_t limit = smem_limit(s->device);
        const int top_cap = range ? 1 : (int)ef + 1;
        // candidate set: room for the live frontier (<= ef + ties) plus stale entries between prunes
        int cand_cap = range ? 4096 : (int)(2 * ef + 64);
        size_t smem = carve_bytes(dt, a.pivot_bytes, a.max_links, top_cap, cand_cap, 0);
        while (smem > limit && cand_cap > 64) {
            cand_cap /= 2;
            smem = carve_bytes(dt, a.pivot_bytes, a.max_links, top_cap, cand_cap, 0);
        }
        if (smem > limit) {
            set_error("hnsw search: ef / dim too large for the shared-memory working set");
            return (int)VSGPU_ERR_ARG;
        }
        a.cand_cap = cand_cap;
        // top-k with ef <= 64: one warp per query (VSGPU_HNSW_CTA=1 forces the CTA-per-query kernel for A/B runs)
        static const bool force_cta = getenv("VSGPU_HNSW_CTA") != nullptr;
        if (!range && !with_spill && ef <= 64 && !force_cta) {
            int sms = 0;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device);
            sms = std::max(sms, 1);
            const int wq_cand = (int)(2 * ef + 64);
            const size_t per_warp = wq_warp_bytes(dt, a.pivot_bytes, a.max_links, wq_cand);
            // enough queries per CTA to fill the machine, few enough that every SM gets work at small batches
            int wpc = (int)std::min<size_t>(WQ_MAX_WARPS, std::max<size_t>(1, (nq + 2 * (size_t)sms - 1) / (2 * (size_t)sms)));
            while (wpc > 1 && (size_t)wpc * per_warp > limit / 2) wpc--;
            if ((size_t)wpc * per_warp <= limit) {
                SearchArgs w = a;
                w.cand_cap = wq_cand;
                const size_t wsmem = (size_t)wpc * per_warp;
                const unsigned grid = (unsigned)((nq + wpc - 1) / wpc);
                if (nq <= 1024) { // latency-bound: wide slices
                    auto kern = hnsw_search_warp_kernel<P, 1>;
                    VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
                    kern<<<grid, wpc * 32, wsmem, s->stream>>>(w, wpc, nq);
                } else {
                    auto kern = hnsw_search_warp_kernel<P, 0>;
                    VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem));
                    kern<<<grid, wpc * 32, wsmem, s->stream>>>(w, wpc, nq);
                }
                VS_CUDA(cudaGetLastError());
                return (int)VSGPU_OK;
            }
        }
        if (range) {
            auto kern = hnsw_range_kernel<P>;
            VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)nq, HNSW_THREADS, smem, s->stream>>>(a);
        } else {
            auto kern = hnsw_search_kernel<P>;
            VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)nq, HNSW_THREADS, smem, s->stream>>>(a);
        }
        VS_CUDA(cudaGetLastError());
        return (int)VSGPU_OK;
    });
}

} // namespace vsgpu

extern "C" {

int vsgpu_hnsw_topk_device(vsgpu_hnsw *g, const void *queries, size_t nq, size_t qstride, size_t k, size_t ef,
                           uint64_t *out_labels, void *out_scores, uint32_t *out_ids, uint32_t *out_counts) {
    vsgpu_store *s = g->s;
    if (nq == 0 || k == 0) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(s->device));
    ef = std::max(ef, k);
    const void *q = nullptr;
    size_t qs = 0;
    const float *qn = nullptr;
    VS_TRY(stage_queries_device(s, queries, nq, qstride, &q, &qs, &qn));
    VS_TRY(ensure_scratch(s, g->misc, nq * 4 + 256));
    uint32_t *status = (uint32_t *)g->misc.ptr;
    VS_CUDA(cudaMemsetAsync(g->counters, 0, 128, s->stream));
    VS_CUDA(cudaEventRecord(s->ev0, s->stream));
    VS_TRY(hnsw_search_core(g, q, nq, qs, qn, k, ef, false, 0, 0, 0, out_ids, out_scores, out_labels, out_counts, nullptr,
                            status, false));
    VS_CUDA(cudaEventRecord(s->ev1, s->stream));
    // overflowed candidate sets (pathological ties / mostly deleted graphs): redo those queries with a spill area
    std::vector<uint32_t> st(nq);
    unsigned long long ctr[10];
    VS_CUDA(cudaMemcpyAsync(st.data(), status, nq * 4, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaMemcpyAsync(ctr, g->counters, 80, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    cudaEventElapsedTime(&g->last_ms, s->ev0, s->ev1);
    g->last_evals = ctr[0];
    g->last_hops = ctr[1];
    if (getenv("VSGPU_HNSW_PROFILE"))
        fprintf(stderr, "[vsgpu_hnsw_topk] nq=%zu ms=%.3f evals=%llu hops=%llu max_hops=%llu cycles/query: gather=%llu eval=%llu admit=%llu pop=%llu bottom=%llu descent=%llu\n",
                nq, g->last_ms, ctr[0], ctr[1], ctr[8], ctr[2] / nq, ctr[3] / nq, ctr[4] / nq, ctr[5] / nq, ctr[6] / nq, ctr[7] / nq);
    const size_t dt = s->type == VSGPU_FLOAT64 ? 8 : 4;
    for (size_t i = 0; i < nq; i++) {
        if (!st[i]) continue;
        VS_TRY(hnsw_search_core(g, (const uint8_t *)q + i * qs, 1, qs, qn ? qn + i : nullptr, k, ef, false, 0, 0, 0,
                                out_ids ? out_ids + i * k : nullptr, out_scores ? (uint8_t *)out_scores + i * k * dt : nullptr,
                                out_labels ? out_labels + i * k : nullptr, out_counts ? out_counts + i : nullptr, nullptr,
                                nullptr, true));
        VS_CUDA(cudaStreamSynchronize(s->stream));
    }
    return VSGPU_OK;
}

int vsgpu_hnsw_topk(vsgpu_hnsw *g, const void *queries, size_t nq, size_t qstride, size_t k, size_t ef,
                    uint64_t *out_labels, double *out_scores, uint32_t *out_ids, uint32_t *out_counts) {
    vsgpu_store *s = g->s;
    if (nq == 0) return VSGPU_OK;
    if (!queries || qstride < s->blob_bytes) {
        set_error("vsgpu_hnsw_topk: bad query buffer");
        return VSGPU_ERR_ARG;
    }
    if (k == 0 || g->count == 0 || g->entry < 0) {
        if (out_counts) std::fill(out_counts, out_counts + nq, 0u);
        return VSGPU_OK;
    }
    VS_CUDA(cudaSetDevice(s->device));
    const size_t dt = s->type == VSGPU_FLOAT64 ? 8 : 4;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t qbytes = nq * s->blob_bytes;
    const size_t o_lab = al(qbytes), o_sc = o_lab + al(nq * k * 8), o_id = o_sc + al(nq * k * dt), o_cnt = o_id + al(nq * k * 4);
    VS_TRY(ensure_pinned(s, o_cnt + al(nq * 4)));
    uint8_t *pin = (uint8_t *)s->pinned;
    for (size_t i = 0; i < nq; i++) memcpy(pin + i * s->blob_bytes, (const uint8_t *)queries + i * qstride, s->blob_bytes);
    VS_TRY(ensure_scratch(s, s->q_raw, al(qbytes)));
    VS_TRY(ensure_scratch(s, g->out, al(nq * k * 8) + al(nq * k * dt) + al(nq * k * 4) + al(nq * 4)));
    uint64_t *d_lab = (uint64_t *)g->out.ptr;
    uint8_t *d_sc = (uint8_t *)g->out.ptr + al(nq * k * 8);
    uint32_t *d_id = (uint32_t *)(d_sc + al(nq * k * dt));
    uint32_t *d_cnt = (uint32_t *)((uint8_t *)d_id + al(nq * k * 4));
    VS_CUDA(cudaMemcpyAsync(s->q_raw.ptr, pin, qbytes, cudaMemcpyHostToDevice, s->stream));
    VS_TRY(vsgpu_hnsw_topk_device(g, s->q_raw.ptr, nq, s->blob_bytes, k, ef, d_lab, d_sc, d_id, d_cnt));
    VS_CUDA(cudaMemcpyAsync(pin + o_lab, d_lab, nq * k * 8, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaMemcpyAsync(pin + o_sc, d_sc, nq * k * dt, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaMemcpyAsync(pin + o_id, d_id, nq * k * 4, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaMemcpyAsync(pin + o_cnt, d_cnt, nq * 4, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    const uint32_t *h_cnt = (const uint32_t *)(pin + o_cnt);
    for (size_t i = 0; i < nq; i++) {
        if (out_counts) out_counts[i] = h_cnt[i];
        for (size_t j = 0; j < k; j++) {
            const size_t o = i * k + j;
            const bool valid = j < h_cnt[i];
            if (out_labels) out_labels[o] = valid ? ((const uint64_t *)(pin + o_lab))[o] : ~0ull;
            if (out_ids) out_ids[o] = valid ? ((const uint32_t *)(pin + o_id))[o] : INV;
            if (out_scores)
                out_scores[o] = !valid ? std::numeric_limits<double>::quiet_NaN()
                                       : (dt == 8 ? ((const double *)(pin + o_sc))[o] : (double)((const float *)(pin + o_sc))[o]);
        }
    }
    return VSGPU_OK;
}

int vsgpu_hnsw_range(vsgpu_hnsw *g, const void *query, double radius, double epsilon, size_t cap, uint64_t *out_labels,
                     double *out_scores, uint32_t *out_ids, size_t *out_count) {
    vsgpu_store *s = g->s;
    if (!query || !out_count) {
        set_error("vsgpu_hnsw_range: bad arguments");
        return VSGPU_ERR_ARG;
    }
    *out_count = 0;
    if (g->count == 0 || g->entry < 0) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(s->device));
    const size_t dt = s->type == VSGPU_FLOAT64 ? 8 : 4;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    VS_TRY(ensure_pinned(s, al(s->blob_bytes)));
    memcpy(s->pinned, query, s->blob_bytes);
    VS_TRY(ensure_scratch(s, s->q_raw, al(s->blob_bytes)));
    VS_CUDA(cudaMemcpyAsync(s->q_raw.ptr, s->pinned, s->blob_bytes, cudaMemcpyHostToDevice, s->stream));
    const void *q = nullptr;
    size_t qs = 0;
    const float *qn = nullptr;
    VS_TRY(stage_queries_device(s, s->q_raw.ptr, 1, s->blob_bytes, &q, &qs, &qn));
    const size_t capd = std::max<size_t>(cap, 1);
    VS_TRY(ensure_scratch(s, g->out, al(capd * 8) + al(capd * dt) + al(capd * 4) + 256));
    uint64_t *d_lab = (uint64_t *)g->out.ptr;
    uint8_t *d_sc = (uint8_t *)g->out.ptr + al(capd * 8);
    uint32_t *d_id = (uint32_t *)(d_sc + al(capd * dt));
    unsigned long long *d_cnt = (unsigned long long *)((uint8_t *)d_id + al(capd * 4));
    uint32_t *d_status = (uint32_t *)(d_cnt + 1);
    VS_CUDA(cudaMemsetAsync(g->counters, 0, 128, s->stream));
    unsigned long long cnt = 0;
    uint32_t st = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        VS_TRY(hnsw_search_core(g, q, 1, qs, qn, 0, 0, true, radius, epsilon, cap, d_id, d_sc, d_lab, nullptr, d_cnt,
                                d_status, attempt == 1));
        VS_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, s->stream));
        VS_CUDA(cudaMemcpyAsync(&st, d_status, 4, cudaMemcpyDeviceToHost, s->stream));
        VS_CUDA(cudaStreamSynchronize(s->stream));
        if (!st) break;
    }
    *out_count = (size_t)cnt;
    if (cnt > cap) return VSGPU_ERR_OVERFLOW;
    if (cnt == 0) return VSGPU_OK;
    if (out_labels) VS_CUDA(cudaMemcpy(out_labels, d_lab, cnt * 8, cudaMemcpyDeviceToHost));
    if (out_ids) VS_CUDA(cudaMemcpy(out_ids, d_id, cnt * 4, cudaMemcpyDeviceToHost));
    if (out_scores) {
        if (dt == 8) {
            VS_CUDA(cudaMemcpy(out_scores, d_sc, cnt * 8, cudaMemcpyDeviceToHost));
        } else {
            std::vector<float> tmp(cnt);
            VS_CUDA(cudaMemcpy(tmp.data(), d_sc, cnt * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < cnt; i++) out_scores[i] = tmp[i];
        }
    }
    return VSGPU_OK;
}

/* ---- resumable batch iterator ---- */
struct vsgpu_hnsw_iter {
    vsgpu_hnsw *g = nullptr;
    size_t cap = 0;     // nodes the state was sized for
    size_t ef = 0;
    uint8_t *query = nullptr; // staged (packed) query on the device
    float q_norm = 0.f;
    uint32_t *visited = nullptr;
    BiState *state = nullptr;
    void *cand_d = nullptr, *extra_d = nullptr, *top_d = nullptr;
    uint32_t *cand_id = nullptr, *extra_id = nullptr, *top_id = nullptr;
    size_t top_cap = 0;
    void *out = nullptr;
    size_t out_cap = 0;
    uint32_t *returned = nullptr;        // multi-value: one bit per row whose label was handed out already (device)
    std::vector<uint32_t> returned_host; // host mirror
    bool returned_dirty = false;
};

static int iter_reset_state(vsgpu_hnsw_iter *it) {
    vsgpu_store *s = it->g->s;
    BiState st{};
    st.entry = -1;
    st.lower = std::numeric_limits<double>::infinity();
    VS_CUDA(cudaMemcpyAsync(it->state, &st, sizeof(st), cudaMemcpyHostToDevice, s->stream));
    VS_CUDA(cudaMemsetAsync(it->visited, 0, ((it->cap + 31) / 32) * 4, s->stream));
    if (it->returned) {
        VS_CUDA(cudaMemsetAsync(it->returned, 0, ((it->cap + 31) / 32) * 4, s->stream));
        std::fill(it->returned_host.begin(), it->returned_host.end(), 0u);
        it->returned_dirty = false;
    }
    VS_CUDA(cudaStreamSynchronize(s->stream));
    return VSGPU_OK;
}

vsgpu_hnsw_iter *vsgpu_hnsw_iter_create(vsgpu_hnsw *g, const void *query, size_t ef) {
    vsgpu_store *s = g->s;
    if (!query || cudaSetDevice(s->device) != cudaSuccess) return nullptr;
    auto *it = new vsgpu_hnsw_iter();
    it->g = g;
    it->cap = std::max<size_t>(g->count, 1);
    it->ef = std::max<size_t>(ef, 1);
    const size_t dt = s->type == VSGPU_FLOAT64 ? 8 : 4;
    bool ok = cudaMalloc(&it->query, s->row_stride + 256) == cudaSuccess;
    ok = ok && cudaMalloc(&it->visited, ((it->cap + 31) / 32) * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&it->state, sizeof(BiState)) == cudaSuccess;
    if (ok && g->multi) {
        ok = cudaMalloc(&it->returned, ((it->cap + 31) / 32) * 4) == cudaSuccess;
        it->returned_host.assign((it->cap + 31) / 32, 0u);
    }
    ok = ok && cudaMalloc(&it->cand_d, (it->cap + 1) * dt) == cudaSuccess && cudaMalloc(&it->cand_id, (it->cap + 1) * 4) == cudaSuccess;
    ok = ok && cudaMalloc(&it->extra_d, (it->cap + 1) * dt) == cudaSuccess && cudaMalloc(&it->extra_id, (it->cap + 1) * 4) == cudaSuccess;
    if (ok) {
        // stage the query: raw blob -> device -> the layout the kernels read
        ok = ensure_pinned(s, s->blob_bytes + 256) == VSGPU_OK && ensure_scratch(s, s->q_raw, s->blob_bytes + 256) == VSGPU_OK;
        if (ok) {
            memcpy(s->pinned, query, s->blob_bytes);
            ok = cudaMemcpyAsync(s->q_raw.ptr, s->pinned, s->blob_bytes, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
        }
        const void *q = nullptr;
        size_t qs = 0;
        const float *qn = nullptr;
        ok = ok && stage_queries_device(s, s->q_raw.ptr, 1, s->blob_bytes, &q, &qs, &qn) == VSGPU_OK;
        if (ok) {
            const size_t bytes = s->plan.kind == CK_INT ? s->row_stride : s->row_bytes;
            ok = cudaMemcpyAsync(it->query, q, bytes, cudaMemcpyDeviceToDevice, s->stream) == cudaSuccess;
            if (ok && qn) ok = cudaMemcpyAsync(&it->q_norm, qn, 4, cudaMemcpyDeviceToHost, s->stream) == cudaSuccess;
            ok = ok && cudaStreamSynchronize(s->stream) == cudaSuccess;
        }
    }
    if (!ok || iter_reset_state(it) != VSGPU_OK) {
        set_error("vsgpu_hnsw_iter_create: device allocation / staging failed");
        vsgpu_hnsw_iter_destroy(it);
        return nullptr;
    }
    return it;
}

void vsgpu_hnsw_iter_destroy(vsgpu_hnsw_iter *it) {
    if (!it) return;
    cudaSetDevice(it->g->s->device);
    cudaStreamSynchronize(it->g->s->stream);
    for (void *p : {(void *)it->query, (void *)it->visited, (void *)it->state, it->cand_d, (void *)it->cand_id, it->extra_d,
                    (void *)it->extra_id, it->top_d, (void *)it->top_id, it->out, (void *)it->returned})
        if (p) cudaFree(p);
    delete it;
}

// multi-value graphs: rows (every row of every label the last batch returned) that later batches must skip
int vsgpu_hnsw_iter_mark_returned(vsgpu_hnsw_iter *it, const uint32_t *ids, size_t n) {
    if (!it->returned) return VSGPU_OK;
    for (size_t i = 0; i < n; i++)
        if (ids[i] < it->cap) it->returned_host[ids[i] >> 5] |= 1u << (ids[i] & 31);
    it->returned_dirty = it->returned_dirty || n > 0;
    return VSGPU_OK;
}

int vsgpu_hnsw_iter_reset(vsgpu_hnsw_iter *it) {
    VS_CUDA(cudaSetDevice(it->g->s->device));
    return iter_reset_state(it);
}

int vsgpu_hnsw_iter_next(vsgpu_hnsw_iter *it, size_t n_res, size_t label_count, uint64_t *out_labels, double *out_scores,
                         uint32_t *out_ids, size_t *out_count, int *depleted) {
    vsgpu_hnsw *g = it->g;
    vsgpu_store *s = g->s;
    VS_CUDA(cudaSetDevice(s->device));
    if (g->count > it->cap) {
        set_error("vsgpu_hnsw_iter_next: the graph grew since the iterator was created");
        return VSGPU_ERR_ARG;
    }
    const size_t dt = s->type == VSGPU_FLOAT64 ? 8 : 4;
    const size_t ef = std::max(it->ef, n_res); // hnsw_batch_iterator.h:210-213
    if (ef + 1 > it->top_cap) {
        if (it->top_d) cudaFree(it->top_d);
        if (it->top_id) cudaFree(it->top_id);
        it->top_d = it->top_id = nullptr;
        VS_CUDA(cudaMalloc(&it->top_d, (ef + 1) * dt));
        VS_CUDA(cudaMalloc(&it->top_id, (ef + 1) * 4));
        it->top_cap = ef + 1;
    }
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t nr = std::max<size_t>(n_res, 1);
    const size_t need = al(nr * 8) + al(nr * dt) + al(nr * 4) + 256;
    if (need > it->out_cap) {
        if (it->out) cudaFree(it->out);
        it->out = nullptr;
        VS_CUDA(cudaMalloc(&it->out, need));
        it->out_cap = need;
    }
    uint64_t *d_lab = (uint64_t *)it->out;
    uint8_t *d_sc = (uint8_t *)it->out + al(nr * 8);
    uint32_t *d_id = (uint32_t *)(d_sc + al(nr * dt));
    uint32_t *d_cnt = (uint32_t *)((uint8_t *)d_id + al(nr * 4));
    BiArgs a{};
    a.k = make_kctx(s);
    a.g = make_graph(g);
    a.q = it->query;
    a.q_norm = it->q_norm;
    a.labels = s->labels;
    a.visited = it->visited;
    a.state = it->state;
    a.cand_d = it->cand_d;
    a.cand_id = it->cand_id;
    a.extra_d = it->extra_d;
    a.extra_id = it->extra_id;
    a.top_d = it->top_d;
    a.top_id = it->top_id;
    a.ef = (int)ef;
    a.n_res = (int)n_res;
    a.max_links = g->M0;
    a.label_count = label_count;
    a.out_ids = d_id;
    a.out_scores = d_sc;
    a.out_labels = d_lab;
    a.out_count = d_cnt;
    a.counters = g->counters;
    a.multi = g->multi ? 1 : 0;
    a.returned_bits = it->returned;
    if (it->returned && it->returned_dirty) {
        VS_CUDA(cudaMemcpyAsync(it->returned, it->returned_host.data(), it->returned_host.size() * 4, cudaMemcpyHostToDevice, s->stream));
        it->returned_dirty = false;
    }
    VS_CUDA(cudaMemsetAsync(g->counters, 0, 128, s->stream));
    const int rc = dispatch_policy(s, [&]<class P>() -> int {
        a.pivot_bytes = P::pivot_bytes(s);
        const size_t smem = carve_bytes(dt, a.pivot_bytes, a.max_links, 1, 1, 0);
        auto kern = hnsw_bi_kernel<P>;
        VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<1, HNSW_THREADS, smem, s->stream>>>(a);
        VS_CUDA(cudaGetLastError());
        return (int)VSGPU_OK;
    });
    VS_TRY(rc);
    uint32_t cnt = 0;
    BiState st{};
    VS_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaMemcpyAsync(&st, it->state, sizeof(st), cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    if (out_count) *out_count = cnt;
    if (depleted) *depleted = (st.depleted && st.extra_n == 0) ? 1 : 0; // isDepleted(), hnsw_batch_iterator.h:251-254
    if (cnt) {
        if (out_labels) VS_CUDA(cudaMemcpy(out_labels, d_lab, cnt * 8, cudaMemcpyDeviceToHost));
        if (out_ids) VS_CUDA(cudaMemcpy(out_ids, d_id, cnt * 4, cudaMemcpyDeviceToHost));
        if (out_scores) {
            if (dt == 8) {
                VS_CUDA(cudaMemcpy(out_scores, d_sc, cnt * 8, cudaMemcpyDeviceToHost));
            } else {
                std::vector<float> tmp(cnt);
                VS_CUDA(cudaMemcpy(tmp.data(), d_sc, cnt * 4, cudaMemcpyDeviceToHost));
                for (size_t i = 0; i < cnt; i++) out_scores[i] = tmp[i];
            }
        }
    }
    return VSGPU_OK;
}

} // extern "C"
