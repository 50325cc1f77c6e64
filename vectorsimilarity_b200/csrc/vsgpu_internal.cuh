// Internal declarations shared by the libvsgpu.so translation units (not part of the C-ABI).
#pragma once
#include "../../include/vsgpu.h"
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace vsgpu {

void set_error(const std::string &msg);

#define VS_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            ::vsgpu::set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " @" +         \
                               __FILE__ + ":" + std::to_string(__LINE__));                         \
            return e_ == cudaErrorMemoryAllocation ? VSGPU_ERR_NOMEM : VSGPU_ERR_CUDA;             \
        }                                                                                          \
    } while (0)

#define VS_TRY(expr)                                                                               \
    do {                                                                                           \
        int rc_ = (expr);                                                                          \
        if (rc_ != VSGPU_OK) return rc_;                                                           \
    } while (0)

// How the reference's CPU kernel for (type, dim, metric) orders its floating-point operations
// (SURVEY App. A3/A7). A "chain" is one SIMD lane of one accumulator register: a
// strictly sequential FMA recurrence. G chains per (row, query) pair, S steps each, then a
// butterfly reduction over the G partial sums — which is exactly what the x86 kernels do.
enum ChainKind : int {
    CK_LANES = 0,      // AVX512F fp32 / fp16 (G=32) and fp64 (G=16): 2 accumulators x L lanes
    CK_BF16_DP = 1,    // AVX512_BF16 vdpbf16ps, bf16 IP (G=16, pairs, odd element first, FTZ)
    CK_BF16_VBMI2 = 2, // AVX512BW+VBMI2 bf16 L2 (G=16, unpacklo then unpackhi)
    CK_SEQ = 3,        // scalar kernels (mul and add rounded separately) and the F16C 8..15 case
    CK_INT = 4         // integer types: exact in any order
};

struct ChainPlan {
    int kind;
    int G;       // chains (threads) per (row, query)
    int S;       // steps per chain
    int dim;
    int r;       // residual = dim % block
    int head;    // r % L
    int nfull;   // r / L
    int prefix;  // number of residual steps before the whole blocks
    int is_l2;
    int ftz;
    int seq_f16c; // CK_SEQ only: the fp16 F16C tier for 8 <= dim < 16
};

ChainPlan make_plan(int type, int metric, size_t dim);

// Element index chain c touches at step s, or -1 for a zero-padded slot.
__host__ __device__ inline int chain_elem(const ChainPlan &p, int c, int s) {
    if (p.kind == CK_LANES) {
        const int L = p.G / 2;
        if (s < p.prefix) {
            if (c < L) return c < p.head ? c : -1;
            return p.nfull ? p.head + (c - L) : -1;
        }
        return p.r + p.G * (s - p.prefix) + c;
    }
    if (p.kind == CK_BF16_DP) {
        // two steps per 32-element block: element 2c+1 first, then 2c
        const int blk = s >> 1, odd = (s & 1) == 0;
        const int within = 2 * c + (odd ? 1 : 0);
        if (p.prefix && blk == 0) return within < p.r ? within : -1;
        return p.r + 32 * (blk - (p.prefix ? 1 : 0)) + within;
    }
    if (p.kind == CK_BF16_VBMI2) {
        int st = s;
        int off = 0;
        if (p.r >= 16) {
            if (st == 0) return c;
            st -= 1;
            off = 16;
        }
        if (p.r % 16) {
            if (st == 0) return c < (p.r % 16) ? off + c : -1;
            st -= 1;
        }
        const int blk = st >> 1, hi = st & 1;
        const int q = c >> 2, i = c & 3;
        return p.r + 32 * blk + 8 * q + (hi ? 4 : 0) + i;
    }
    return s;
}

struct Scratch {
    void *ptr = nullptr;
    size_t bytes = 0;
};

// A tensor-path top-k whose overflow flags have not been looked at yet. The call itself never waits for the device:
// the per-query flags (candidate buffer overflowed: adversarial ties inside the error band) are copied to pinned memory
// behind the kernels, and whoever synchronises next (vsgpu_store_sync, the host-buffer vsgpu_topk, the next call on the
// store) redoes the flagged queries on the exact path into the same output buffers.
struct PendingTopk {
    bool active = false;
    const void *q = nullptr; // staged queries (device) as the kernels read them
    size_t nq = 0, q_stride = 0, k = 0;
    const float *q_norms = nullptr;
    uint32_t *out_ids = nullptr;
    void *out_scores = nullptr;
    uint64_t *out_labels = nullptr;
    size_t n_events = 0;  // (start, stop) event pairs recorded around the scan launches
    size_t n_totals = 0;  // candidate counters (one per chunk of MAX_NQ queries)
};

} // namespace vsgpu

struct vsgpu_store {
    int device = 0;
    int type = 0, metric = 0;
    size_t dim = 0;
    size_t elem = 0;        // bytes per element
    size_t row_bytes = 0;   // dim * elem
    size_t row_stride = 0;  // row_bytes rounded up to 16 (TMA / vector loads)
    size_t blob_bytes = 0;  // what the host sees per row (row_bytes + 4 for int8/uint8 cosine)
    bool has_norm = false;  // int8/uint8 cosine
    size_t count = 0, capacity = 0;
    uint8_t *rows = nullptr;
    uint64_t *labels = nullptr;
    float *norms = nullptr;   // has_norm: the reference's appended norm, one per row
    // tensor-path mirrors (allocated lazily)
    uint16_t *shadow = nullptr; // fp32 stores: bf16 (RNE) copy of the rows, stride = shadow_stride elems
    size_t shadow_stride = 0;
    float *row_l2 = nullptr;    // ||row||_2 (fp32, rounded up) for the coarse-pass error bound
    float max_row_l2 = 0.f;
    bool shadow_valid_upto_count = false;
    cudaStream_t stream = nullptr;
    bool own_stream = true; // false: adopted from the caller (vsgpu_store_set_stream), never destroyed here
    vsgpu::ChainPlan plan{};
    // scratch (grown on demand, owned by the store)
    vsgpu::Scratch q_raw, q_dev, scores, sel_state, out_dev, cand, misc;
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    vsgpu_stats stats{};
    vsgpu::PendingTopk pend;
    vsgpu::Scratch ovf;          // device: [nq] u32 overflow flags of the last tensor-path call, then u64 candidate totals
    uint32_t *h_ovf = nullptr;   // pinned mirror of `ovf`
    size_t h_ovf_bytes = 0;
    std::vector<cudaEvent_t> scan_evs;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    void *tmap_cache = nullptr; // tensor path: cached tensor maps
};

namespace vsgpu {

int ensure_scratch(vsgpu_store *s, Scratch &sc, size_t bytes);
int ensure_pinned(vsgpu_store *s, size_t bytes);
// tensor path: device flag / counter block for nq queries in `chunks` chunks (zeroed), and its pinned mirror
int pending_begin(vsgpu_store *s, size_t nq, size_t chunks, uint32_t **d_flags, unsigned long long **d_totals);
// enqueue the copy of flags + totals to pinned memory and arm s->pend
int pending_arm(vsgpu_store *s, const void *q, size_t nq, size_t q_stride, const float *q_norms, size_t k, uint32_t *out_ids,
                void *out_scores, uint64_t *out_labels, size_t n_events, size_t chunks);
cudaEvent_t scan_event(vsgpu_store *s, size_t i); // i-th reusable event of the store (created on demand)
// waits for the stream, redoes overflowed queries exactly, completes the stats; no-op when nothing is pending
int resolve_pending_topk(vsgpu_store *s);
// raw query blobs on the device -> what the kernels read (zero-padded rows + norms for integer types)
int stage_queries_device(vsgpu_store *s, const void *q_dev_raw, size_t nq, size_t qstride, const void **q_out,
                         size_t *q_stride_out, const float **q_norms_out);

// ---- exact path (vsgpu_exact.cu) ----
// scores[q * ld + id] for q < nq, id < n: DistType (float, or double for fp64 stores).
int launch_exact_scan(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms,
                      void *scores, size_t ld);
// scores for chosen ids: out[q * ld + i] = dist(row ids[q * ids_ld + i], query q); id = UINT32_MAX -> +inf/NaN slot
int launch_exact_gather(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms,
                        const uint32_t *ids, size_t ids_ld, const uint32_t *counts, size_t max_count,
                        void *out, size_t ld);

// ---- staged streaming scan (vsgpu_scan.cu): fp32 / fp16 stores, dim % 32 == 0, <= 16 raw queries per pass ----
bool tma_scan_supported(const vsgpu_store *s);
int launch_tma_scan(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, void *scores, size_t ld);
// scan + selection in one pass (per-CTA running top-k in shared memory + one merge launch): <= 8 queries, k <= 128
bool fused_topk_supported(const vsgpu_store *s, size_t nq, size_t k);
int launch_fused_topk(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, size_t k, size_t out_ld, uint32_t *out_ids,
                      void *out_scores, uint64_t *out_labels);

// ---- selection (vsgpu_select.cu) ----
// For each of nq score rows (DistType, length n, leading dim ld): the k smallest by (score, id),
// sorted, into out_ids/out_scores ([nq][k]) and labels gathered from s->labels.
int launch_select_topk(vsgpu_store *s, const void *scores, size_t ld, size_t nq, size_t n, size_t k, size_t out_ld,
                       uint32_t *out_ids, void *out_scores, uint64_t *out_labels);
// Sort per-query candidate lists (ids + exact scores) and keep k: lists of length counts[q] <= max_count.
int launch_sort_candidates(vsgpu_store *s, size_t nq, size_t k, const uint32_t *cand_ids, const void *cand_scores,
                           size_t cand_ld, const uint32_t *counts, size_t out_ld, uint32_t *out_ids, void *out_scores,
                           uint64_t *out_labels);
size_t select_sort_max();
int launch_range_compact(vsgpu_store *s, const void *scores, size_t n, double radius, size_t cap,
                         uint32_t *out_ids, void *out_scores, uint64_t *out_labels, unsigned long long *out_count);

// ---- tensor path (vsgpu_tensor.cu) ----
bool tensor_path_supported(const vsgpu_store *s, size_t nq, size_t k);
// phased call of a sharded index (vsgpu_topk_device_begin / _next / _finish): `world` shards run `rounds` coarse phases
// each and exchange `bounds` ([2 nq] fp32, device) after every one
struct PhasedCall {
    unsigned world, rounds;
    float *bounds;
};
int tensor_topk(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms, size_t k,
                uint32_t *out_ids, void *out_scores, uint64_t *out_labels, const PhasedCall *ph = nullptr);
void tensor_topk_reset(vsgpu_store *s);
int tensor_topk_next(vsgpu_store *s, float *bounds);
int tensor_topk_finish(vsgpu_store *s, const float *bounds);
size_t tensor_topk_rounds(size_t rows, size_t k, unsigned world);
int tensor_sync_mirrors(vsgpu_store *s);
void tensor_release(vsgpu_store *s);
// a single row changed (src != SIZE_MAX: row src was copied over row id): patch the tensor-path side data in place
int tensor_row_changed(vsgpu_store *s, size_t id, size_t src);
int tensor_i8_row_changed(vsgpu_store *s, size_t id, size_t src);
// int8 / uint8 stores: exact integer GEMM on tcgen05 kind::i8 (vsgpu_tensor_i8.cu)
bool tensor_i8_supported(const vsgpu_store *s, size_t nq, size_t k);
int tensor_i8_topk(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms, size_t k,
                   uint32_t *out_ids, void *out_scores, uint64_t *out_labels);
void tensor_i8_release(vsgpu_store *s);

} // namespace vsgpu
