// Selection kernels: k smallest (score, internal id) per query out of a dense score row (MSB-first
// radix select, ties at the k-th score resolved by ascending id exactly like a sequential scan with
// strict-< admission, reference algorithms/brute_force/brute_force.h:262-288), block bitonic sort
// of the survivors, range compaction (brute_force.h:304-321) and the k-way merge of shard lists.
// All integer / compare work, bounded by HBM traffic over the 4- or 8-byte scores.
#include "vsgpu_internal.cuh"
#include <algorithm>

namespace vsgpu {

// order-preserving score -> unsigned key (NaN sorts last, as a row that never beats the bound)
__device__ __forceinline__ uint32_t to_key(float v) {
    uint32_t u = __float_as_uint(v);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t to_key(double v) {
    uint64_t u = (uint64_t)__double_as_longlong(v);
    if ((u & 0x7fffffffffffffffull) > 0x7ff0000000000000ull) return ~0ull;
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ float from_key(uint32_t k) {
    if (k == 0xffffffffu) return __uint_as_float(0x7fc00000u);
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ double from_key(uint64_t k) {
    if (k == ~0ull) return __longlong_as_double(0x7ff8000000000000ll);
    return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
}

template <typename KT> struct SelState {
    KT prefix;              // key bits decided so far
    uint32_t k_rem;         // rank still to locate inside the current prefix bucket
    uint32_t less_ctr;      // append cursor for keys < T
    uint32_t pad;
    unsigned hist[256];
};

constexpr int SEG = 2048;
__host__ __device__ inline size_t zmin(size_t a, size_t b) { return a < b ? a : b; }
// // elements per warp segment (contiguous, keeps ties in id order)

template <typename ST, typename KT>
__global__ void __launch_bounds__(256) radix_hist_kernel(const ST *__restrict__ scores, size_t ld, size_t n, int pass,
                                                         SelState<KT> *__restrict__ st) {
    __shared__ unsigned h[256];
    const int q = blockIdx.y;
    h[threadIdx.x] = 0;
    __syncthreads();
    constexpr int BITS = sizeof(KT) * 8;
    const int shift = BITS - 8 * (pass + 1);
    const KT prefix = st[q].prefix;
    const KT mask_hi = pass == 0 ? KT(0) : (~KT(0)) << (shift + 8);
    const ST *row = scores + (size_t)q * ld;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const KT key = to_key(row[i]);
        if ((key & mask_hi) == prefix) atomicAdd(&h[(unsigned)((key >> shift) & 0xff)], 1u);
    }
    __syncthreads();
    const unsigned v = h[threadIdx.x];
    if (v) atomicAdd(&st[q].hist[threadIdx.x], v);
}

template <typename KT>
__global__ void radix_pick_kernel2(SelState<KT> *__restrict__ st, int nq, int pass) {
    const int q = blockIdx.x;
    if (q >= nq) return;
    __shared__ unsigned h[256];
    h[threadIdx.x] = st[q].hist[threadIdx.x];
    st[q].hist[threadIdx.x] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        constexpr int BITS = sizeof(KT) * 8;
        const int shift = BITS - 8 * (pass + 1);
        unsigned k = st[q].k_rem, cum = 0;
        int d = 0;
        for (; d < 256; d++) {
            if (cum + h[d] >= k) break;
            cum += h[d];
        }
        if (d == 256) d = 255; // cannot happen when k <= n
        st[q].k_rem = k - cum;
        st[q].prefix |= (KT)d << shift;
    }
}

template <typename KT>
__global__ void sel_init_kernel(SelState<KT> *st, int nq, uint32_t k) {
    const int q = blockIdx.x;
    if (q >= nq) return;
    st[q].hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        st[q].prefix = 0;
        st[q].k_rem = k;
        st[q].less_ctr = 0;
    }
}

template <typename ST, typename KT>
__global__ void __launch_bounds__(256) tie_count_kernel(const ST *__restrict__ scores, size_t ld, size_t n, size_t nseg,
                                                        const SelState<KT> *__restrict__ st, unsigned *__restrict__ tie_cnt) {
    const int q = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    const KT T = st[q].prefix;
    const ST *row = scores + (size_t)q * ld;
    for (size_t seg = warp; seg < nseg; seg += nwarps) {
        const size_t b = seg * SEG, e = zmin(n, b + SEG);
        unsigned cnt = 0;
        for (size_t i = b + lane; i < e; i += 32) cnt += to_key(row[i]) == T;
        for (int w = 16; w >= 1; w >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, w);
        if (lane == 0) tie_cnt[(size_t)q * nseg + seg] = cnt;
    }
}

// exclusive scan of tie_cnt[q][0..nseg) in place
__global__ void __launch_bounds__(1024) tie_scan_kernel(unsigned *__restrict__ tie_cnt, size_t nseg) {
    __shared__ unsigned part[1024];
    unsigned *row = tie_cnt + (size_t)blockIdx.x * nseg;
    const size_t per = (nseg + blockDim.x - 1) / blockDim.x;
    const size_t b = zmin(threadIdx.x * per, nseg), e = zmin(b + per, nseg);
    unsigned sum = 0;
    for (size_t i = b; i < e; i++) sum += row[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
        unsigned v = threadIdx.x >= (unsigned)off ? part[threadIdx.x - off] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = part[threadIdx.x] - sum;
    for (size_t i = b; i < e; i++) {
        const unsigned v = row[i];
        row[i] = run;
        run += v;
    }
}

template <typename ST, typename KT>
__global__ void __launch_bounds__(256) sel_compact_kernel(const ST *__restrict__ scores, size_t ld, size_t n, size_t nseg,
                                                          SelState<KT> *__restrict__ st, const unsigned *__restrict__ tie_base,
                                                          uint32_t k, KT *__restrict__ out_keys, uint32_t *__restrict__ out_ids) {
    const int q = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const size_t warp = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    const KT T = st[q].prefix;
    const uint32_t m = st[q].k_rem;  // ties to take
    const uint32_t less = k - m;     // keys strictly below T
    const ST *row = scores + (size_t)q * ld;
    KT *ok = out_keys + (size_t)q * k;
    uint32_t *oi = out_ids + (size_t)q * k;
    for (size_t seg = warp; seg < nseg; seg += nwarps) {
        const size_t b = seg * SEG, e = zmin(n, b + SEG);
        unsigned tie_run = tie_base[(size_t)q * nseg + seg];
        for (size_t i0 = b; i0 < e; i0 += 32) {
            const size_t i = i0 + lane;
            const bool in = i < e;
            const KT key = in ? to_key(row[i]) : ~KT(0);
            const bool is_less = in && key < T;
            const bool is_tie = in && key == T;
            const unsigned bl = __ballot_sync(0xffffffffu, is_less);
            const unsigned bt = __ballot_sync(0xffffffffu, is_tie);
            if (bl) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(&st[q].less_ctr, (unsigned)__popc(bl));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (is_less) {
                    const unsigned pos = base + __popc(bl & ((1u << lane) - 1));
                    if (pos < less) { ok[pos] = key; oi[pos] = (uint32_t)i; }
                }
            }
            if (bt) {
                if (is_tie) {
                    const unsigned rank = tie_run + __popc(bt & ((1u << lane) - 1));
                    if (rank < m) { ok[less + rank] = key; oi[less + rank] = (uint32_t)i; }
                }
                tie_run += __popc(bt);
            }
        }
    }
}

// Bitonic sort of up to P (power of two) (key, id) pairs per block, ascending (key, id); writes the
// first k to the outputs and gathers labels.
template <typename ST, typename KT>
__global__ void __launch_bounds__(1024) sort_pairs_kernel(const KT *__restrict__ in_keys, const ST *__restrict__ in_scores,
                                                          const uint32_t *__restrict__ in_ids, size_t in_ld,
                                                          const uint32_t *__restrict__ counts, uint32_t fixed_count, int P_max,
                                                          uint32_t k, uint32_t out_ld, const uint64_t *__restrict__ labels,
                                                          uint32_t *__restrict__ out_ids, ST *__restrict__ out_scores,
                                                          uint64_t *__restrict__ out_labels) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KT *sk = reinterpret_cast<KT *>(smem_raw);
    uint32_t *si = reinterpret_cast<uint32_t *>(smem_raw + (size_t)P_max * sizeof(KT));
    const int q = blockIdx.x;
    uint32_t cnt = counts ? counts[q] : fixed_count;
    if (cnt > (uint32_t)P_max) cnt = P_max;
    // the network is sized for THIS list (block-uniform): a pruned candidate list of a few dozen entries does not pay for
    // the P_max = 1024-wide sort its buffer could hold
    int P = 2; // <= P_max: cnt <= P_max and P_max is a power of two
    while (P < (int)cnt) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        KT key = ~KT(0);
        uint32_t id = 0xffffffffu;
        if ((uint32_t)i < cnt) {
            id = in_ids[(size_t)q * in_ld + i];
            key = in_keys ? in_keys[(size_t)q * in_ld + i] : to_key(in_scores[(size_t)q * in_ld + i]);
            if (id == 0xffffffffu) key = ~KT(0);
        }
        sk[i] = key;
        si[i] = id;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < P / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const KT ka = sk[lo], kb = sk[hi];
                const uint32_t ia = si[lo], ib = si[hi];
                const bool a_gt_b = ka > kb || (ka == kb && ia > ib);
                if (a_gt_b == asc) {
                    sk[lo] = kb; sk[hi] = ka;
                    si[lo] = ib; si[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < out_ld; i += blockDim.x) {
        const bool ok = i < k && i < cnt && si[i] != 0xffffffffu;
        const uint32_t id = ok ? si[i] : 0xffffffffu;
        if (out_ids) out_ids[(size_t)q * out_ld + i] = id;
        if (out_scores) out_scores[(size_t)q * out_ld + i] = ok ? from_key(sk[i]) : from_key(~KT(0));
        if (out_labels) out_labels[(size_t)q * out_ld + i] = ok ? labels[id] : ~0ull;
    }
}

// Small problems (one or a few queries over <= a few hundred thousand rows) are bound by launch latency, not bandwidth:
// the radix select above is 13 launches. Here each block sorts one 2048-score chunk by (key, id) in shared memory and
// hands its k best to the final block sort: 2 launches, same (score, id) order and tie rule.
constexpr int CHUNK = 2048;
template <typename ST, typename KT>
__global__ void __launch_bounds__(256) chunk_topk_kernel(const ST *__restrict__ scores, size_t ld, size_t n, uint32_t k,
                                                         KT *__restrict__ out_keys, uint32_t *__restrict__ out_ids, size_t out_ld) {
    __shared__ KT sk[CHUNK];
    __shared__ uint32_t si[CHUNK];
    const int q = blockIdx.y;
    const size_t first = (size_t)blockIdx.x * CHUNK;
    const ST *row = scores + (size_t)q * ld;
    for (int i = threadIdx.x; i < CHUNK; i += blockDim.x) {
        const size_t r = first + i;
        sk[i] = r < n ? to_key(row[r]) : ~KT(0);
        si[i] = r < n ? (uint32_t)r : 0xffffffffu;
    }
    __syncthreads();
    for (int size = 2; size <= CHUNK; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < CHUNK / 2; t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool asc = (lo & size) == 0;
                const KT ka = sk[lo], kb = sk[hi];
                const uint32_t ia = si[lo], ib = si[hi];
                const bool a_gt_b = ka > kb || (ka == kb && ia > ib);
                if (a_gt_b == asc) {
                    sk[lo] = kb; sk[hi] = ka;
                    si[lo] = ib; si[hi] = ia;
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        const size_t o = (size_t)q * out_ld + (size_t)blockIdx.x * k + i;
        out_keys[o] = sk[i];
        out_ids[o] = si[i];
    }
}

// unsorted finalisation for k beyond the in-block sort limit (the host sorts)
template <typename ST, typename KT>
__global__ void finalize_unsorted_kernel(const KT *__restrict__ keys, const uint32_t *__restrict__ ids, size_t total,
                                         const uint64_t *__restrict__ labels, ST *__restrict__ out_scores,
                                         uint64_t *__restrict__ out_labels) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        if (out_scores) out_scores[i] = from_key(keys[i]);
        if (out_labels) out_labels[i] = labels[ids[i]];
    }
}

template <typename ST>
__global__ void __launch_bounds__(256) range_compact_kernel(const ST *__restrict__ scores, size_t n, ST radius, size_t cap,
                                                            const uint64_t *__restrict__ labels, uint32_t *__restrict__ out_ids,
                                                            ST *__restrict__ out_scores, uint64_t *__restrict__ out_labels,
                                                            unsigned long long *__restrict__ counter) {
    const int lane = threadIdx.x & 31;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n_round = (n + 31) / 32 * 32;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        const bool in = i < n;
        const ST v = in ? scores[i] : ST(0);
        const bool hit = in && v <= radius; // brute_force.h:316 (NaN never matches)
        const unsigned b = __ballot_sync(0xffffffffu, hit);
        if (b) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(counter, (unsigned long long)__popc(b));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) {
                const unsigned long long pos = base + __popc(b & ((1u << lane) - 1));
                if (pos < cap) {
                    out_ids[pos] = (uint32_t)i;
                    out_scores[pos] = v;
                    out_labels[pos] = labels[i];
                }
            }
        }
    }
}

// k-way merge of `parts` sorted lists per query by ascending (score, label): each output slot
// counts how many entries precede it (lists are short: parts*k <= a few thousand).
template <typename ST>
__global__ void merge_lists_kernel(size_t parts, size_t nq, size_t k, const ST *__restrict__ scores,
                                   const uint64_t *__restrict__ labels, ST *__restrict__ out_scores,
                                   uint64_t *__restrict__ out_labels) {
    const size_t q = blockIdx.x;
    const size_t total = parts * k;
    for (size_t e = threadIdx.x; e < total; e += blockDim.x) {
        const size_t p = e / k, j = e % k;
        const size_t src = (p * nq + q) * k + j;
        const uint64_t lab = labels[src];
        if (lab == ~0ull) continue;
        const auto key = to_key(scores[src]);
        size_t rank = 0;
        for (size_t p2 = 0; p2 < parts; p2++) {
            const ST *sc2 = scores + (p2 * nq + q) * k;
            const uint64_t *lb2 = labels + (p2 * nq + q) * k;
            // entries of list p2 that sort before (key, lab): binary search on the sorted list
            size_t lo = 0, hi = k;
            while (lo < hi) {
                const size_t mid = (lo + hi) / 2;
                const uint64_t l2 = lb2[mid];
                bool before;
                if (l2 == ~0ull) before = false;
                else {
                    const auto k2 = to_key(sc2[mid]);
                    before = k2 < key || (k2 == key && (l2 < lab || (l2 == lab && p2 < p)));
                }
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_scores[q * k + rank] = scores[src];
            out_labels[q * k + rank] = lab;
        }
    }
}

template <typename ST>
__global__ void fill_pad_kernel(ST *scores, uint64_t *labels, size_t total) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        scores[i] = from_key(~decltype(to_key(ST(0)))(0));
        labels[i] = ~0ull;
    }
}

// ------------------------------------------------------------------------------------------------
static int next_pow2(size_t v) {
    int p = 1;
    while ((size_t)p < v) p <<= 1;
    return p;
}

constexpr size_t SORT_MAX = 4096;

template <typename ST, typename KT>
static int select_topk_t(vsgpu_store *s, const void *scores_v, size_t ld, size_t nq, size_t n, size_t k, size_t out_ld,
                         uint32_t *out_ids, void *out_scores, uint64_t *out_labels) {
    const ST *scores = (const ST *)scores_v;
    {
        // launch-latency-bound shapes: chunk sort + final sort (2 launches instead of 13)
        const size_t chunks = (n + CHUNK - 1) / CHUNK;
        if (nq <= 8 && k <= CHUNK && chunks * k <= 4096 && chunks <= 1024) {
            const size_t cand = chunks * k;
            auto al2 = [](size_t v) { return (v + 255) / 256 * 256; };
            VS_TRY(ensure_scratch(s, s->sel_state, al2(sizeof(KT) * nq * cand) + al2(sizeof(uint32_t) * nq * cand)));
            auto *ck = (KT *)s->sel_state.ptr;
            auto *ci = (uint32_t *)((unsigned char *)s->sel_state.ptr + al2(sizeof(KT) * nq * cand));
            chunk_topk_kernel<ST, KT><<<dim3((unsigned)chunks, (unsigned)nq), 256, 0, s->stream>>>(scores, ld, n, (uint32_t)k, ck, ci, cand);
            const int P = next_pow2(cand);
            const size_t smem = (size_t)P * (sizeof(KT) + sizeof(uint32_t));
            auto kern = sort_pairs_kernel<ST, KT>;
            if (smem > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int threads = std::max(32, std::min(1024, P / 2));
            kern<<<(unsigned)nq, threads, smem, s->stream>>>(ck, nullptr, ci, cand, nullptr, (uint32_t)cand, P, (uint32_t)k, (uint32_t)out_ld,
                                                            s->labels, out_ids, (ST *)out_scores, out_labels);
            VS_CUDA(cudaGetLastError());
            s->stats.kernel_launches += 2;
            return VSGPU_OK;
        }
    }
    const size_t nseg = (n + SEG - 1) / SEG;
    // scratch: states | tie counts | keys | ids
    const size_t st_bytes = sizeof(SelState<KT>) * nq;
    const size_t tie_bytes = sizeof(unsigned) * nq * nseg;
    const size_t key_bytes = sizeof(KT) * nq * k;
    const size_t id_bytes = sizeof(uint32_t) * nq * k;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    VS_TRY(ensure_scratch(s, s->sel_state, al(st_bytes) + al(tie_bytes) + al(key_bytes) + al(id_bytes)));
    unsigned char *base = (unsigned char *)s->sel_state.ptr;
    auto *st = (SelState<KT> *)base;
    auto *tie = (unsigned *)(base + al(st_bytes));
    auto *keys = (KT *)(base + al(st_bytes) + al(tie_bytes));
    auto *ids = (uint32_t *)(base + al(st_bytes) + al(tie_bytes) + al(key_bytes));
    cudaStream_t str = s->stream;

    sel_init_kernel<KT><<<(unsigned)nq, 256, 0, str>>>(st, (int)nq, (uint32_t)k);
    s->stats.kernel_launches++;
    unsigned bx = (unsigned)std::min<size_t>((n + 256 * 16 - 1) / (256 * 16), 1184);
    if (bx == 0) bx = 1;
    // keep the whole grid around a few waves when many queries are selected at once
    while (bx > 8 && (size_t)bx * nq > 1184 * 8) bx /= 2;
    constexpr int PASSES = sizeof(KT);
    for (int pass = 0; pass < PASSES; pass++) {
        radix_hist_kernel<ST, KT><<<dim3(bx, (unsigned)nq), 256, 0, str>>>(scores, ld, n, pass, st);
        radix_pick_kernel2<KT><<<(unsigned)nq, 256, 0, str>>>(st, (int)nq, pass);
        s->stats.kernel_launches += 2;
    }
    unsigned wb = (unsigned)std::min<size_t>((nseg + 7) / 8, 1184);
    if (wb == 0) wb = 1;
    while (wb > 8 && (size_t)wb * nq > 1184 * 8) wb /= 2;
    tie_count_kernel<ST, KT><<<dim3(wb, (unsigned)nq), 256, 0, str>>>(scores, ld, n, nseg, st, tie);
    tie_scan_kernel<<<(unsigned)nq, 1024, 0, str>>>(tie, nseg);
    sel_compact_kernel<ST, KT><<<dim3(wb, (unsigned)nq), 256, 0, str>>>(scores, ld, n, nseg, st, tie, (uint32_t)k, keys, ids);
    s->stats.kernel_launches += 3;
    if (k <= SORT_MAX) {
        const int P = next_pow2(k);
        const size_t smem = (size_t)P * (sizeof(KT) + sizeof(uint32_t));
        auto kern = sort_pairs_kernel<ST, KT>;
        if (smem > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int threads = std::max(32, std::min(1024, P / 2));
        kern<<<(unsigned)nq, threads, smem, str>>>(keys, nullptr, ids, k, nullptr, (uint32_t)k, P, (uint32_t)k, (uint32_t)out_ld, s->labels,
                                                  out_ids, (ST *)out_scores, out_labels);
    } else {
        // too long for one block: hand back unsorted, the C-ABI layer sorts on the host
        if (out_ld != k) {
            set_error("select: k beyond the device sort limit needs out_ld == k");
            return VSGPU_ERR_ARG;
        }
        if (out_ids) VS_CUDA(cudaMemcpyAsync(out_ids, ids, id_bytes, cudaMemcpyDeviceToDevice, str));
        finalize_unsorted_kernel<ST, KT><<<256, 256, 0, str>>>(keys, ids, nq * k, s->labels, (ST *)out_scores, out_labels);
    }
    s->stats.kernel_launches++;
    VS_CUDA(cudaGetLastError());
    return VSGPU_OK;
}

int launch_select_topk(vsgpu_store *s, const void *scores, size_t ld, size_t nq, size_t n, size_t k, size_t out_ld,
                       uint32_t *out_ids, void *out_scores, uint64_t *out_labels) {
    if (nq == 0 || k == 0 || n == 0) return VSGPU_OK;
    if (k > n) {
        set_error("select: k > n (caller clamps)");
        return VSGPU_ERR_ARG;
    }
    if (s->type == VSGPU_FLOAT64)
        return select_topk_t<double, uint64_t>(s, scores, ld, nq, n, k, out_ld, out_ids, out_scores, out_labels);
    return select_topk_t<float, uint32_t>(s, scores, ld, nq, n, k, out_ld, out_ids, out_scores, out_labels);
}

size_t select_sort_max() { return SORT_MAX; }

template <typename ST, typename KT>
static int sort_candidates_t(vsgpu_store *s, size_t nq, size_t k, const uint32_t *cand_ids, const void *cand_scores,
                             size_t cand_ld, const uint32_t *counts, size_t out_ld, uint32_t *out_ids, void *out_scores,
                             uint64_t *out_labels) {
    const int P = next_pow2(std::max(cand_ld, k));
    const size_t smem = (size_t)P * (sizeof(KT) + sizeof(uint32_t));
    auto kern = sort_pairs_kernel<ST, KT>;
    if (smem > 48 * 1024) VS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int threads = std::max(32, std::min(1024, P / 2));
    kern<<<(unsigned)nq, threads, smem, s->stream>>>(nullptr, (const ST *)cand_scores, cand_ids, cand_ld, counts,
                                                    (uint32_t)cand_ld, P, (uint32_t)k, (uint32_t)out_ld, s->labels, out_ids,
                                                    (ST *)out_scores, out_labels);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

int launch_sort_candidates(vsgpu_store *s, size_t nq, size_t k, const uint32_t *cand_ids, const void *cand_scores,
                           size_t cand_ld, const uint32_t *counts, size_t out_ld, uint32_t *out_ids, void *out_scores,
                           uint64_t *out_labels) {
    if (nq == 0 || k == 0) return VSGPU_OK;
    if (s->type == VSGPU_FLOAT64)
        return sort_candidates_t<double, uint64_t>(s, nq, k, cand_ids, cand_scores, cand_ld, counts, out_ld, out_ids, out_scores, out_labels);
    return sort_candidates_t<float, uint32_t>(s, nq, k, cand_ids, cand_scores, cand_ld, counts, out_ld, out_ids, out_scores, out_labels);
}

int launch_range_compact(vsgpu_store *s, const void *scores, size_t n, double radius, size_t cap, uint32_t *out_ids,
                         void *out_scores, uint64_t *out_labels, unsigned long long *out_count) {
    VS_CUDA(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), s->stream));
    if (n == 0) return VSGPU_OK;
    unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 1184);
    if (s->type == VSGPU_FLOAT64)
        range_compact_kernel<double><<<blocks, 256, 0, s->stream>>>((const double *)scores, n, radius, cap, s->labels, out_ids,
                                                                    (double *)out_scores, out_labels, out_count);
    else
        range_compact_kernel<float><<<blocks, 256, 0, s->stream>>>((const float *)scores, n, (float)radius, cap, s->labels,
                                                                   out_ids, (float *)out_scores, out_labels, out_count);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

} // namespace vsgpu

extern "C" int vsgpu_merge_topk_device(int device, void *stream, int dtype_f64, size_t parts, size_t nq, size_t k,
                                       const void *scores, const uint64_t *labels, void *out_scores,
                                       uint64_t *out_labels) {
    using namespace vsgpu;
    if (parts == 0 || nq == 0 || k == 0) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(device));
    cudaStream_t str = (cudaStream_t)stream;
    const unsigned threads = (unsigned)std::min<size_t>(1024, std::max<size_t>(32, (parts * k + 31) / 32 * 32));
    if (dtype_f64) {
        fill_pad_kernel<double><<<64, 256, 0, str>>>((double *)out_scores, out_labels, nq * k);
        merge_lists_kernel<double><<<(unsigned)nq, threads, 0, str>>>(parts, nq, k, (const double *)scores, labels,
                                                                      (double *)out_scores, out_labels);
    } else {
        fill_pad_kernel<float><<<64, 256, 0, str>>>((float *)out_scores, out_labels, nq * k);
        merge_lists_kernel<float><<<(unsigned)nq, threads, 0, str>>>(parts, nq, k, (const float *)scores, labels,
                                                                     (float *)out_scores, out_labels);
    }
    VS_CUDA(cudaGetLastError());
    return VSGPU_OK;
}
