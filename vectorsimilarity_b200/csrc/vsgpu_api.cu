// The C-ABI of libvsgpu.so (include/vsgpu.h): device store lifecycle and the orchestration of the
// exact / tensor query paths. Host logic only; kernels live in vsgpu_exact.cu, vsgpu_select.cu,
// vsgpu_tensor.cu.
#include "vsgpu_internal.cuh"
#include <algorithm>
#include <cmath>
#include <limits>
#include <mutex>
#include <numeric>
#include <vector>

namespace vsgpu {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

int ensure_scratch(vsgpu_store *s, Scratch &sc, size_t bytes) {
    if (bytes <= sc.bytes) return VSGPU_OK;
    if (sc.ptr) {
        VS_CUDA(cudaStreamSynchronize(s->stream));
        VS_CUDA(cudaFree(sc.ptr));
        sc.ptr = nullptr;
        sc.bytes = 0;
    }
    size_t want = std::max(bytes, (size_t)1 << 16);
    want = (want + 255) / 256 * 256;
    VS_CUDA(cudaMalloc(&sc.ptr, want));
    sc.bytes = want;
    return VSGPU_OK;
}

int ensure_pinned(vsgpu_store *s, size_t bytes) {
    if (bytes <= s->pinned_bytes) return VSGPU_OK;
    if (s->pinned) {
        VS_CUDA(cudaStreamSynchronize(s->stream));
        VS_CUDA(cudaFreeHost(s->pinned));
        s->pinned = nullptr;
        s->pinned_bytes = 0;
    }
    size_t want = std::max(bytes, (size_t)1 << 16);
    VS_CUDA(cudaMallocHost(&s->pinned, want));
    s->pinned_bytes = want;
    return VSGPU_OK;
}

int pending_begin(vsgpu_store *s, size_t nq, size_t chunks, uint32_t **d_flags, unsigned long long **d_totals) {
    const size_t fl = (nq * 4 + 255) / 256 * 256, bytes = fl + chunks * 8;
    VS_TRY(ensure_scratch(s, s->ovf, bytes));
    if (s->h_ovf_bytes < bytes) {
        if (s->h_ovf) {
            VS_CUDA(cudaStreamSynchronize(s->stream));
            VS_CUDA(cudaFreeHost(s->h_ovf));
            s->h_ovf = nullptr;
            s->h_ovf_bytes = 0;
        }
        const size_t want = std::max<size_t>(bytes, 32768);
        VS_CUDA(cudaMallocHost(&s->h_ovf, want));
        s->h_ovf_bytes = want;
    }
    VS_CUDA(cudaMemsetAsync(s->ovf.ptr, 0, bytes, s->stream));
    *d_flags = (uint32_t *)s->ovf.ptr;
    *d_totals = (unsigned long long *)((uint8_t *)s->ovf.ptr + fl);
    return VSGPU_OK;
}

int pending_arm(vsgpu_store *s, const void *q, size_t nq, size_t q_stride, const float *q_norms, size_t k, uint32_t *out_ids,
                void *out_scores, uint64_t *out_labels, size_t n_events, size_t chunks) {
    const size_t fl = (nq * 4 + 255) / 256 * 256;
    VS_CUDA(cudaMemcpyAsync(s->h_ovf, s->ovf.ptr, fl + chunks * 8, cudaMemcpyDeviceToHost, s->stream));
    PendingTopk &p = s->pend;
    p.active = true;
    p.q = q;
    p.nq = nq;
    p.q_stride = q_stride;
    p.q_norms = q_norms;
    p.k = k;
    p.out_ids = out_ids;
    p.out_scores = out_scores;
    p.out_labels = out_labels;
    p.n_events = n_events;
    p.n_totals = chunks;
    return VSGPU_OK;
}

cudaEvent_t scan_event(vsgpu_store *s, size_t i) {
    while (s->scan_evs.size() <= i) {
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        s->scan_evs.push_back(e);
    }
    return s->scan_evs[i];
}

int resolve_pending_topk(vsgpu_store *s) {
    if (!s->pend.active) return VSGPU_OK;
    PendingTopk p = s->pend;
    s->pend.active = false;
    VS_CUDA(cudaStreamSynchronize(s->stream));
    const size_t fl = (p.nq * 4 + 255) / 256 * 256;
    const auto *tot = (const unsigned long long *)((const uint8_t *)s->h_ovf + fl);
    for (size_t c = 0; c < p.n_totals; c++) s->stats.candidates += tot[c];
    for (size_t e = 0; e < p.n_events; e++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, s->scan_evs[2 * e], s->scan_evs[2 * e + 1]) == cudaSuccess) s->stats.scan_ms += ms;
    }
    const size_t n = s->count, ld = (n + 63) / 64 * 64;
    const size_t ssz = s->type == VSGPU_FLOAT64 ? 8 : 4;
    bool any = false;
    for (size_t q = 0; q < p.nq; q++) {
        if (!s->h_ovf[q]) continue;
        any = true;
        s->stats.fallback_queries++;
        VS_TRY(ensure_scratch(s, s->scores, ld * ssz));
        VS_TRY(launch_exact_scan(s, (const uint8_t *)p.q + q * p.q_stride, 1, p.q_stride, p.q_norms ? p.q_norms + q : nullptr,
                                 s->scores.ptr, ld));
        VS_TRY(launch_select_topk(s, s->scores.ptr, ld, 1, n, p.k, p.k, p.out_ids ? p.out_ids + q * p.k : nullptr,
                                  p.out_scores ? (uint8_t *)p.out_scores + q * p.k * ssz : nullptr,
                                  p.out_labels ? p.out_labels + q * p.k : nullptr));
    }
    if (any) VS_CUDA(cudaStreamSynchronize(s->stream));
    return VSGPU_OK;
}

static size_t elem_size(int type) {
    switch (type) {
    case VSGPU_FLOAT32: return 4;
    case VSGPU_FLOAT64: return 8;
    case VSGPU_BFLOAT16:
    case VSGPU_FLOAT16: return 2;
    default: return 1;
    }
}

static int grow(vsgpu_store *s, size_t need) {
    if (need <= s->capacity) return VSGPU_OK;
    size_t cap = std::max<size_t>(need, s->capacity + s->capacity / 2);
    cap = std::max<size_t>(cap, 1024);
    if (cap > 0xfffffffeull) {
        set_error("store capacity exceeds idType range");
        return VSGPU_ERR_ARG;
    }
    uint8_t *rows = nullptr;
    uint64_t *labels = nullptr;
    float *norms = nullptr;
    VS_CUDA(cudaMalloc(&rows, cap * s->row_stride));
    VS_CUDA(cudaMalloc(&labels, cap * sizeof(uint64_t)));
    if (s->has_norm) VS_CUDA(cudaMalloc(&norms, cap * sizeof(float)));
    if (s->count) {
        VS_CUDA(cudaMemcpyAsync(rows, s->rows, s->count * s->row_stride, cudaMemcpyDeviceToDevice, s->stream));
        VS_CUDA(cudaMemcpyAsync(labels, s->labels, s->count * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s->stream));
        if (s->has_norm)
            VS_CUDA(cudaMemcpyAsync(norms, s->norms, s->count * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    }
    VS_CUDA(cudaStreamSynchronize(s->stream));
    if (s->rows) cudaFree(s->rows);
    if (s->labels) cudaFree(s->labels);
    if (s->norms) cudaFree(s->norms);
    s->rows = rows;
    s->labels = labels;
    s->norms = norms;
    s->capacity = cap;
    tensor_release(s); // mirrors are rebuilt lazily for the new capacity
    return VSGPU_OK;
}

// raw query blobs (device, `src_stride` apart; int8/uint8 cosine blobs carry the fp32 norm after the
// dim bytes) -> zero-padded rows at the store's row stride + a norm array
__global__ void repack_queries_kernel(const uint8_t *__restrict__ src, size_t src_stride, size_t row_bytes,
                                      size_t row_stride, int has_norm, size_t nq, uint8_t *__restrict__ dst,
                                      float *__restrict__ norms) {
    const size_t total = nq * row_stride;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t q = i / row_stride, b = i % row_stride;
        dst[i] = b < row_bytes ? src[q * src_stride + b] : 0;
    }
    if (has_norm) {
        for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x) {
            const uint8_t *p = src + q * src_stride + row_bytes;
            const unsigned u = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16) | ((unsigned)p[3] << 24);
            norms[q] = __uint_as_float(u);
        }
    }
}

__global__ void int_norm_kernel(const uint8_t *__restrict__ rows, size_t row_stride, size_t dim, int is_unsigned, size_t n,
                                float *__restrict__ norms) {
    // float(sqrt(double(uint64 sum x^2))) — spaces/normalize/compute_norm.h:17-31
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint8_t *p = rows + i * row_stride;
        unsigned long long sum = 0;
        for (size_t e = 0; e < dim; e++) {
            const int v = is_unsigned ? (int)p[e] : (int)(signed char)p[e];
            sum += (unsigned long long)(v * v);
        }
        norms[i] = __double2float_rn(sqrt((double)sum));
    }
}

__global__ void iota_kernel(uint64_t *p, uint64_t first, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = first + i;
}

template <typename T> __global__ void fill_kernel(T *p, T v, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

int stage_queries_device(vsgpu_store *s, const void *q_dev_raw, size_t nq, size_t qstride, const void **q_out,
                                size_t *q_stride_out, const float **q_norms_out) {
    const bool need_repack = s->plan.kind == CK_INT || s->has_norm;
    if (!need_repack) {
        *q_out = q_dev_raw;
        *q_stride_out = qstride;
        *q_norms_out = nullptr;
        return VSGPU_OK;
    }
    const size_t qbytes = nq * s->row_stride;
    const size_t off_norm = (qbytes + 255) / 256 * 256;
    VS_TRY(ensure_scratch(s, s->q_dev, off_norm + nq * sizeof(float) + 256));
    // layout of q_dev: [packed rows] [norms]
    uint8_t *packed = (uint8_t *)s->q_dev.ptr;
    float *norms = (float *)((uint8_t *)s->q_dev.ptr + off_norm);
    unsigned blocks = (unsigned)std::min<size_t>((qbytes + 255) / 256, 2048);
    repack_queries_kernel<<<std::max(blocks, 1u), 256, 0, s->stream>>>((const uint8_t *)q_dev_raw, qstride, s->row_bytes,
                                                                      s->row_stride, s->has_norm ? 1 : 0, nq, packed, norms);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    *q_out = packed;
    *q_stride_out = s->row_stride;
    *q_norms_out = s->has_norm ? norms : nullptr;
    return VSGPU_OK;
}

static size_t score_size(const vsgpu_store *s) { return s->type == VSGPU_FLOAT64 ? 8 : 4; }

// queries staged on the device -> device outputs [nq][out_ld]
static int topk_core(vsgpu_store *s, const void *q_dev, size_t nq, size_t q_stride, const float *q_norms, size_t k,
                     size_t out_ld, unsigned flags, uint32_t *out_ids, void *out_scores, uint64_t *out_labels,
                     const PhasedCall *ph = nullptr) {
    const size_t n = s->count;
    const size_t k_eff = std::min(k, n);
    const size_t ssz = score_size(s);
    if (k_eff == 0) {
        // nothing to select: pad
        const size_t total = nq * out_ld;
        if (total) {
            if (out_ids) fill_kernel<uint32_t><<<64, 256, 0, s->stream>>>(out_ids, 0xffffffffu, total);
            if (out_labels) fill_kernel<uint64_t><<<64, 256, 0, s->stream>>>(out_labels, ~0ull, total);
            if (out_scores) {
                if (ssz == 8) fill_kernel<double><<<64, 256, 0, s->stream>>>((double *)out_scores, std::numeric_limits<double>::quiet_NaN(), total);
                else fill_kernel<float><<<64, 256, 0, s->stream>>>((float *)out_scores, std::numeric_limits<float>::quiet_NaN(), total);
            }
            VS_CUDA(cudaGetLastError());
        }
        return VSGPU_OK;
    }
    const bool want_tensor = flags != VSGPU_TOPK_EXACT_ONLY && tensor_path_supported(s, nq, k_eff);
    if (flags == VSGPU_TOPK_TENSOR_ONLY && !want_tensor) {
        set_error("tensor path not available for this store / batch shape");
        return VSGPU_ERR_ARG;
    }
    if (want_tensor) {
        if (out_ld != k_eff) {
            set_error("tensor path needs k <= size");
            return VSGPU_ERR_ARG;
        }
        s->stats.path = 1;
        return tensor_topk(s, q_dev, nq, q_stride, q_norms, k_eff, out_ids, out_scores, out_labels, ph);
    }
    s->stats.path = 0;
    // small batches on the staged scan: selection fused into the scan (per-CTA running top-k in shared memory) — no score
    // matrix, 2 launches per pass of up to 8 queries instead of 1 + 13
    if (fused_topk_supported(s, std::min<size_t>(nq, 8), k_eff) && nq <= 8) {
        VS_CUDA(cudaEventRecord(s->ev2, s->stream));
        VS_TRY(launch_fused_topk(s, q_dev, nq, q_stride, k_eff, out_ld, out_ids, out_scores, out_labels));
        VS_CUDA(cudaEventRecord(s->ev3, s->stream));
        return VSGPU_OK;
    }
    // exact path: chunks of queries sized so the score matrix stays within ~1/16 of HBM or 2 GB
    const size_t ld = (n + 63) / 64 * 64;
    // one launch = one pass over the store for up to 16 queries (the kernel's register tile)
    size_t budget = (size_t)2 << 30;
    size_t qc = std::max<size_t>(1, std::min<size_t>(std::min<size_t>(nq, 16), budget / (ld * ssz)));
    VS_TRY(ensure_scratch(s, s->scores, qc * ld * ssz));
    for (size_t q0 = 0; q0 < nq; q0 += qc) {
        const size_t nqc = std::min(qc, nq - q0);
        const uint8_t *qp = (const uint8_t *)q_dev + q0 * q_stride;
        VS_CUDA(cudaEventRecord(s->ev2, s->stream));
        VS_TRY(launch_exact_scan(s, qp, nqc, q_stride, q_norms ? q_norms + q0 : nullptr, s->scores.ptr, ld));
        VS_CUDA(cudaEventRecord(s->ev3, s->stream));
        VS_TRY(launch_select_topk(s, s->scores.ptr, ld, nqc, n, k_eff, out_ld, out_ids ? out_ids + q0 * out_ld : nullptr,
                                  out_scores ? (uint8_t *)out_scores + q0 * out_ld * ssz : nullptr,
                                  out_labels ? out_labels + q0 * out_ld : nullptr));
    }
    return VSGPU_OK;
}

} // namespace vsgpu

using namespace vsgpu;

extern "C" {

const char *vsgpu_last_error(void) { return g_err.c_str(); }
void vsgpu_set_last_error(const char *msg) { g_err = msg ? msg : ""; }

int vsgpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int vsgpu_mem_info(int device, size_t *free_bytes, size_t *total_bytes) {
    VS_CUDA(cudaSetDevice(device));
    VS_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
    return VSGPU_OK;
}

vsgpu_store *vsgpu_store_create(int device, int type, int metric, size_t dim, size_t capacity_hint) {
    if (type < 0 || type > VSGPU_UINT8 || metric < 0 || metric > VSGPU_COSINE || dim == 0 || dim > (1u << 24)) {
        set_error("vsgpu_store_create: bad type/metric/dim");
        return nullptr;
    }
    int ndev = vsgpu_device_count();
    if (device < 0 || device >= ndev) {
        set_error("vsgpu_store_create: no such CUDA device (libvsgpu needs a GPU; there is no CPU fallback)");
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("cudaSetDevice failed");
        return nullptr;
    }
    auto *s = new vsgpu_store();
    s->device = device;
    s->type = type;
    s->metric = metric;
    s->dim = dim;
    s->elem = elem_size(type);
    s->row_bytes = dim * s->elem;
    s->row_stride = (s->row_bytes + 15) / 16 * 16;
    s->has_norm = metric == VSGPU_COSINE && (type == VSGPU_INT8 || type == VSGPU_UINT8);
    s->blob_bytes = s->row_bytes + (s->has_norm ? 4 : 0);
    s->plan = make_plan(type, metric, dim);
    bool ok = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&s->ev0) == cudaSuccess && cudaEventCreate(&s->ev1) == cudaSuccess;
    ok = ok && cudaEventCreate(&s->ev2) == cudaSuccess && cudaEventCreate(&s->ev3) == cudaSuccess;
    if (ok && capacity_hint) ok = grow(s, capacity_hint) == VSGPU_OK;
    if (!ok) {
        if (g_err.empty()) set_error("vsgpu_store_create: CUDA resource creation failed");
        vsgpu_store_destroy(s);
        return nullptr;
    }
    return s;
}

void vsgpu_store_destroy(vsgpu_store *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    tensor_release(s);
    for (Scratch *sc : {&s->q_raw, &s->q_dev, &s->scores, &s->sel_state, &s->out_dev, &s->cand, &s->misc})
        if (sc->ptr) cudaFree(sc->ptr);
    if (s->pinned) cudaFreeHost(s->pinned);
    if (s->h_ovf) cudaFreeHost(s->h_ovf);
    if (s->ovf.ptr) cudaFree(s->ovf.ptr);
    for (cudaEvent_t e : s->scan_evs) cudaEventDestroy(e);
    if (s->rows) cudaFree(s->rows);
    if (s->labels) cudaFree(s->labels);
    if (s->norms) cudaFree(s->norms);
    for (cudaEvent_t e : {s->ev0, s->ev1, s->ev2, s->ev3})
        if (e) cudaEventDestroy(e);
    if (s->stream && s->own_stream) cudaStreamDestroy(s->stream);
    delete s;
}

int vsgpu_store_set_stream(vsgpu_store *s, void *stream) {
    if (!s || !stream) {
        set_error("vsgpu_store_set_stream: bad arguments");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    if (s->own_stream) cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)stream;
    s->own_stream = false;
    return VSGPU_OK;
}

size_t vsgpu_store_size(const vsgpu_store *s) { return s->count; }
size_t vsgpu_store_row_bytes(const vsgpu_store *s) { return s->blob_bytes; }
size_t vsgpu_store_device_bytes(const vsgpu_store *s) {
    size_t b = s->capacity * (s->row_stride + sizeof(uint64_t) + (s->has_norm ? 4 : 0));
    if (s->shadow) b += s->capacity * s->shadow_stride * 2;
    if (s->row_l2) b += s->capacity * 4;
    for (const Scratch *sc : {&s->q_raw, &s->q_dev, &s->scores, &s->sel_state, &s->out_dev, &s->cand, &s->misc}) b += sc->bytes;
    return b;
}
void *vsgpu_store_stream(vsgpu_store *s) { return (void *)s->stream; }
int vsgpu_store_sync(vsgpu_store *s) {
    VS_CUDA(cudaSetDevice(s->device));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    return resolve_pending_topk(s); // a tensor-path top-k enqueued earlier: redo its overflowed queries now
}

int vsgpu_store_append(vsgpu_store *s, const void *rows, size_t stride, const uint64_t *labels, size_t n) {
    if (n == 0) return VSGPU_OK;
    if (!rows || !labels || stride < s->blob_bytes) {
        set_error("vsgpu_store_append: bad arguments");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    VS_TRY(grow(s, s->count + n));
    uint8_t *dst = s->rows + s->count * s->row_stride;
    if (s->row_stride != s->row_bytes) VS_CUDA(cudaMemsetAsync(dst, 0, n * s->row_stride, s->stream));
    VS_CUDA(cudaMemcpy2DAsync(dst, s->row_stride, rows, stride, s->row_bytes, n, cudaMemcpyHostToDevice, s->stream));
    if (s->has_norm)
        VS_CUDA(cudaMemcpy2DAsync(s->norms + s->count, sizeof(float), (const uint8_t *)rows + s->row_bytes, stride,
                                  sizeof(float), n, cudaMemcpyHostToDevice, s->stream));
    VS_CUDA(cudaMemcpyAsync(s->labels + s->count, labels, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream)); // the caller may reuse its buffers
    s->count += n;
    s->shadow_valid_upto_count = false;
    return VSGPU_OK;
}

int vsgpu_store_append_device(vsgpu_store *s, const void *rows, size_t stride, const uint64_t *labels,
                              uint64_t first_label, const float *norms, size_t n) {
    if (n == 0) return VSGPU_OK;
    if (!rows || stride < s->row_bytes) {
        set_error("vsgpu_store_append_device: bad arguments");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    VS_TRY(grow(s, s->count + n));
    uint8_t *dst = s->rows + s->count * s->row_stride;
    if (s->row_stride != s->row_bytes) VS_CUDA(cudaMemsetAsync(dst, 0, n * s->row_stride, s->stream));
    VS_CUDA(cudaMemcpy2DAsync(dst, s->row_stride, rows, stride, s->row_bytes, n, cudaMemcpyDeviceToDevice, s->stream));
    if (labels) {
        VS_CUDA(cudaMemcpyAsync(s->labels + s->count, labels, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, s->stream));
    } else {
        iota_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 4096), 256, 0, s->stream>>>(s->labels + s->count, first_label, n);
        VS_CUDA(cudaGetLastError());
    }
    if (s->has_norm) {
        if (norms) {
            VS_CUDA(cudaMemcpyAsync(s->norms + s->count, norms, n * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
        } else {
            unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 4096);
            int_norm_kernel<<<blocks, 256, 0, s->stream>>>(dst, s->row_stride, s->dim, s->type == VSGPU_UINT8, n,
                                                          s->norms + s->count);
            VS_CUDA(cudaGetLastError());
        }
    }
    VS_CUDA(cudaStreamSynchronize(s->stream));
    s->count += n;
    s->shadow_valid_upto_count = false;
    return VSGPU_OK;
}

int vsgpu_store_update(vsgpu_store *s, size_t id, const void *row, uint64_t label) {
    if (id >= s->count || !row) {
        set_error("vsgpu_store_update: bad id");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    VS_CUDA(cudaMemcpyAsync(s->rows + id * s->row_stride, row, s->row_bytes, cudaMemcpyHostToDevice, s->stream));
    if (s->has_norm)
        VS_CUDA(cudaMemcpyAsync(s->norms + id, (const uint8_t *)row + s->row_bytes, 4, cudaMemcpyHostToDevice, s->stream));
    VS_CUDA(cudaMemcpyAsync(s->labels + id, &label, sizeof(label), cudaMemcpyHostToDevice, s->stream));
    VS_TRY(tensor_row_changed(s, id, (size_t)-1)); // the row's mirror / norms follow; nothing else is rebuilt
    VS_CUDA(cudaStreamSynchronize(s->stream));
    return VSGPU_OK;
}

int vsgpu_store_remove_swap(vsgpu_store *s, size_t dst) {
    if (dst >= s->count) {
        set_error("vsgpu_store_remove_swap: bad id");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    const size_t last = s->count - 1;
    if (dst != last) {
        VS_CUDA(cudaMemcpyAsync(s->rows + dst * s->row_stride, s->rows + last * s->row_stride, s->row_stride,
                                cudaMemcpyDeviceToDevice, s->stream));
        VS_CUDA(cudaMemcpyAsync(s->labels + dst, s->labels + last, sizeof(uint64_t), cudaMemcpyDeviceToDevice, s->stream));
        if (s->has_norm)
            VS_CUDA(cudaMemcpyAsync(s->norms + dst, s->norms + last, sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    }
    s->count = last;
    // the moved row's mirror / norms move with it (the bf16 mirror of a 10 M x 768 store is 15 GB: r1 rebuilt all of it)
    if (dst != last) VS_TRY(tensor_row_changed(s, dst, last));
    else VS_TRY(tensor_row_changed(s, (size_t)-1, (size_t)-1)); // only clamps the mirrored count
    VS_CUDA(cudaStreamSynchronize(s->stream));
    return VSGPU_OK;
}

int vsgpu_store_truncate(vsgpu_store *s, size_t new_count) {
    if (new_count > s->count) {
        set_error("vsgpu_store_truncate: new_count exceeds size");
        return VSGPU_ERR_ARG;
    }
    if (new_count == s->count) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    s->count = new_count;
    s->shadow_valid_upto_count = false;
    tensor_release(s);
    return VSGPU_OK;
}

int vsgpu_store_read(const vsgpu_store *s, size_t first, size_t n, void *rows, size_t stride, uint64_t *labels) {
    if (first + n > s->count) {
        set_error("vsgpu_store_read: range out of bounds");
        return VSGPU_ERR_ARG;
    }
    if (n == 0) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(s->device));
    if (rows) {
        VS_CUDA(cudaMemcpy2DAsync(rows, stride, s->rows + first * s->row_stride, s->row_stride, s->row_bytes, n,
                                  cudaMemcpyDeviceToHost, s->stream));
        if (s->has_norm)
            VS_CUDA(cudaMemcpy2DAsync((uint8_t *)rows + s->row_bytes, stride, s->norms + first, sizeof(float), sizeof(float),
                                      n, cudaMemcpyDeviceToHost, s->stream));
    }
    if (labels)
        VS_CUDA(cudaMemcpyAsync(labels, s->labels + first, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    return VSGPU_OK;
}

int vsgpu_topk_device(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t k, unsigned flags,
                      uint64_t *out_labels, void *out_scores, uint32_t *out_ids) {
    if (nq == 0 || k == 0) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s)); // an earlier call the caller never synchronised on
    s->stats = vsgpu_stats{};
    VS_CUDA(cudaEventRecord(s->ev0, s->stream));
    const void *q = nullptr;
    size_t qs = 0;
    const float *qn = nullptr;
    VS_TRY(stage_queries_device(s, queries, nq, qstride, &q, &qs, &qn));
    VS_TRY(topk_core(s, q, nq, qs, qn, k, k, flags, out_ids, out_scores, out_labels));
    VS_CUDA(cudaEventRecord(s->ev1, s->stream));
    return VSGPU_OK;
}

// Phased variant for sharded callers (`world` shards, one store each). A tensor-path call runs its coarse pass
// in `rounds` phases — _begin the first, every _next one more — and after each writes this shard's bounds to `bounds`
// ([2 nq] fp32, DEVICE): bounds[q] bounds its k-th best score from below, -bounds[nq + q] its ceil(k / world)-th best. The
// caller reduces the whole buffer with MAX over the shards between the calls; max(bounds[q], -bounds[nq + q]) is then a lower
// bound of the k-th best score overall, which every shard admits against in its next phase and prunes against in _finish
// before the exact re-rank. Calls that take another path do all their work in _begin and leave neutral bounds.
int vsgpu_topk_device_begin(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t k, unsigned flags,
                            uint64_t *out_labels, void *out_scores, uint32_t *out_ids, unsigned world, unsigned rounds,
                            float *bounds) {
    if (nq == 0 || k == 0) return VSGPU_OK;
    if (!bounds || world < 2 || rounds == 0) return vsgpu_topk_device(s, queries, nq, qstride, k, flags, out_labels, out_scores, out_ids);
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    s->stats = vsgpu_stats{};
    VS_CUDA(cudaEventRecord(s->ev0, s->stream));
    fill_kernel<float><<<16, 256, 0, s->stream>>>(bounds, -std::numeric_limits<float>::infinity(), nq);
    fill_kernel<float><<<16, 256, 0, s->stream>>>(bounds + nq, std::numeric_limits<float>::infinity(), nq);
    const void *q = nullptr;
    size_t qs = 0;
    const float *qn = nullptr;
    VS_TRY(stage_queries_device(s, queries, nq, qstride, &q, &qs, &qn));
    tensor_topk_reset(s);
    const PhasedCall ph{world, rounds, bounds};
    VS_TRY(topk_core(s, q, nq, qs, qn, k, k, flags, out_ids, out_scores, out_labels, &ph));
    VS_CUDA(cudaEventRecord(s->ev1, s->stream));
    return VSGPU_OK;
}

int vsgpu_topk_device_next(vsgpu_store *s, float *bounds) {
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(tensor_topk_next(s, bounds));
    VS_CUDA(cudaEventRecord(s->ev1, s->stream));
    return VSGPU_OK;
}

int vsgpu_topk_device_finish(vsgpu_store *s, const float *bounds) {
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(tensor_topk_finish(s, bounds));
    VS_CUDA(cudaEventRecord(s->ev1, s->stream));
    return VSGPU_OK;
}

size_t vsgpu_topk_rounds(size_t rows_per_shard, size_t k, unsigned world) { return tensor_topk_rounds(rows_per_shard, k, world); }

int vsgpu_topk(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t k, unsigned flags,
               uint64_t *out_labels, double *out_scores, uint32_t *out_ids, uint32_t *out_counts) {
    if (nq == 0) return VSGPU_OK;
    if (!queries || qstride < s->blob_bytes) {
        set_error("vsgpu_topk: bad query buffer");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    s->stats = vsgpu_stats{};
    const size_t n = s->count;
    const size_t k_eff = std::min(k, n);
    if (out_counts)
        for (size_t q = 0; q < nq; q++) out_counts[q] = (uint32_t)k_eff;
    auto pad_from = [&](size_t from) {
        for (size_t q = 0; q < nq; q++)
            for (size_t j = from; j < k; j++) {
                if (out_labels) out_labels[q * k + j] = ~0ull;
                if (out_scores) out_scores[q * k + j] = std::numeric_limits<double>::quiet_NaN();
                if (out_ids) out_ids[q * k + j] = 0xffffffffu;
            }
    };
    if (k_eff == 0) {
        pad_from(0);
        return VSGPU_OK;
    }
    const size_t ssz = score_size(s);
    // pinned staging: [queries][labels][scores][ids]
    const size_t qbytes = nq * s->blob_bytes;
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_lab = al(qbytes), o_sc = o_lab + al(nq * k_eff * 8), o_id = o_sc + al(nq * k_eff * ssz);
    const size_t pin_total = o_id + al(nq * k_eff * 4);
    VS_TRY(ensure_pinned(s, pin_total));
    uint8_t *pin = (uint8_t *)s->pinned;
    for (size_t q = 0; q < nq; q++) memcpy(pin + q * s->blob_bytes, (const uint8_t *)queries + q * qstride, s->blob_bytes);
    VS_TRY(ensure_scratch(s, s->q_raw, al(qbytes)));
    uint8_t *raw_dev = (uint8_t *)s->q_raw.ptr;
    VS_TRY(ensure_scratch(s, s->out_dev, al(nq * k_eff * 8) + al(nq * k_eff * ssz) + al(nq * k_eff * 4)));
    uint64_t *d_lab = (uint64_t *)s->out_dev.ptr;
    uint8_t *d_sc = (uint8_t *)s->out_dev.ptr + al(nq * k_eff * 8);
    uint32_t *d_id = (uint32_t *)(d_sc + al(nq * k_eff * ssz));

    VS_CUDA(cudaEventRecord(s->ev0, s->stream));
    VS_CUDA(cudaMemcpyAsync(raw_dev, pin, qbytes, cudaMemcpyHostToDevice, s->stream));
    const void *q = nullptr;
    size_t qs = 0;
    const float *qn = nullptr;
    VS_TRY(stage_queries_device(s, raw_dev, nq, s->blob_bytes, &q, &qs, &qn));
    VS_TRY(topk_core(s, q, nq, qs, qn, k_eff, k_eff, flags, d_id, d_sc, d_lab));
    VS_TRY(resolve_pending_topk(s)); // tensor path: overflowed queries are redone before the results leave
    // labels | scores | ids sit at the same 256-byte-aligned offsets on both sides: one copy (a single query is bound by
    // the number of stream operations, not by bytes)
    VS_CUDA(cudaMemcpyAsync(pin + o_lab, d_lab, al(nq * k_eff * 8) + al(nq * k_eff * ssz) + nq * k_eff * 4, cudaMemcpyDeviceToHost,
                            s->stream));
    VS_CUDA(cudaEventRecord(s->ev1, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) s->stats.total_ms = ms;
    if (s->stats.path == 0 && cudaEventElapsedTime(&ms, s->ev2, s->ev3) == cudaSuccess) s->stats.scan_ms = ms;

    const uint64_t *h_lab = (const uint64_t *)(pin + o_lab);
    const uint32_t *h_id = (const uint32_t *)(pin + o_id);
    const bool host_sort = s->stats.path == 0 && k_eff > select_sort_max();
    std::vector<uint32_t> perm;
    for (size_t q = 0; q < nq; q++) {
        const size_t base = q * k_eff;
        auto score_at = [&](size_t i) -> double {
            return ssz == 8 ? ((const double *)(pin + o_sc))[base + i] : (double)((const float *)(pin + o_sc))[base + i];
        };
        if (host_sort) {
            perm.resize(k_eff);
            std::iota(perm.begin(), perm.end(), 0u);
            std::sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) {
                const double sa = score_at(a), sb = score_at(b);
                if (sa != sb) return sa < sb || (sb != sb && sa == sa);
                return h_id[base + a] < h_id[base + b];
            });
        }
        for (size_t j = 0; j < k_eff; j++) {
            const size_t i = host_sort ? perm[j] : j;
            if (out_labels) out_labels[q * k + j] = h_lab[base + i];
            if (out_scores) out_scores[q * k + j] = score_at(i);
            if (out_ids) out_ids[q * k + j] = h_id[base + i];
        }
    }
    pad_from(k_eff);
    return VSGPU_OK;
}

static int one_query_scores(vsgpu_store *s, const void *query, void **scores_dev) {
    // stage the single query and run the exact scan over all rows
    const size_t ssz = score_size(s);
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    VS_TRY(ensure_pinned(s, al(s->blob_bytes)));
    memcpy(s->pinned, query, s->blob_bytes);
    VS_TRY(ensure_scratch(s, s->q_raw, al(s->blob_bytes)));
    uint8_t *raw_dev = (uint8_t *)s->q_raw.ptr;
    VS_CUDA(cudaMemcpyAsync(raw_dev, s->pinned, s->blob_bytes, cudaMemcpyHostToDevice, s->stream));
    const void *q = nullptr;
    size_t qs = 0;
    const float *qn = nullptr;
    VS_TRY(stage_queries_device(s, raw_dev, 1, s->blob_bytes, &q, &qs, &qn));
    const size_t ld = (s->count + 63) / 64 * 64;
    VS_TRY(ensure_scratch(s, s->scores, std::max<size_t>(ld, 64) * ssz));
    VS_TRY(launch_exact_scan(s, q, 1, qs, qn, s->scores.ptr, ld));
    *scores_dev = s->scores.ptr;
    return VSGPU_OK;
}

int vsgpu_range(vsgpu_store *s, const void *query, double radius, size_t cap, uint64_t *out_labels, double *out_scores,
                uint32_t *out_ids, size_t *out_count) {
    if (!query || !out_count) {
        set_error("vsgpu_range: bad arguments");
        return VSGPU_ERR_ARG;
    }
    *out_count = 0;
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    s->stats = vsgpu_stats{};
    if (s->count == 0) return VSGPU_OK;
    const size_t ssz = score_size(s);
    void *scores = nullptr;
    VS_TRY(one_query_scores(s, query, &scores));
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t capd = std::max<size_t>(cap, 1);
    VS_TRY(ensure_scratch(s, s->out_dev, al(capd * 8) + al(capd * ssz) + al(capd * 4) + 256));
    uint64_t *d_lab = (uint64_t *)s->out_dev.ptr;
    uint8_t *d_sc = (uint8_t *)s->out_dev.ptr + al(capd * 8);
    uint32_t *d_id = (uint32_t *)(d_sc + al(capd * ssz));
    unsigned long long *d_cnt = (unsigned long long *)((uint8_t *)d_id + al(capd * 4));
    VS_TRY(launch_range_compact(s, scores, s->count, radius, cap, d_id, d_sc, d_lab, d_cnt));
    unsigned long long cnt = 0;
    VS_CUDA(cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, s->stream));
    VS_CUDA(cudaStreamSynchronize(s->stream));
    *out_count = (size_t)cnt;
    if (cnt > cap) return VSGPU_ERR_OVERFLOW;
    if (cnt == 0) return VSGPU_OK;
    if (out_labels) VS_CUDA(cudaMemcpy(out_labels, d_lab, cnt * 8, cudaMemcpyDeviceToHost));
    if (out_ids) VS_CUDA(cudaMemcpy(out_ids, d_id, cnt * 4, cudaMemcpyDeviceToHost));
    if (out_scores) {
        if (ssz == 8) {
            VS_CUDA(cudaMemcpy(out_scores, d_sc, cnt * 8, cudaMemcpyDeviceToHost));
        } else {
            std::vector<float> tmp(cnt);
            VS_CUDA(cudaMemcpy(tmp.data(), d_sc, cnt * 4, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < cnt; i++) out_scores[i] = tmp[i];
        }
    }
    return VSGPU_OK;
}

int vsgpu_scores(vsgpu_store *s, const void *query, double *out_scores) {
    if (!query || !out_scores) {
        set_error("vsgpu_scores: bad arguments");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    s->stats = vsgpu_stats{};
    if (s->count == 0) return VSGPU_OK;
    void *scores = nullptr;
    VS_TRY(one_query_scores(s, query, &scores));
    if (score_size(s) == 8) {
        VS_CUDA(cudaMemcpyAsync(out_scores, scores, s->count * 8, cudaMemcpyDeviceToHost, s->stream));
        VS_CUDA(cudaStreamSynchronize(s->stream));
    } else {
        std::vector<float> tmp(s->count);
        VS_CUDA(cudaMemcpyAsync(tmp.data(), scores, s->count * 4, cudaMemcpyDeviceToHost, s->stream));
        VS_CUDA(cudaStreamSynchronize(s->stream));
        for (size_t i = 0; i < s->count; i++) out_scores[i] = tmp[i];
    }
    return VSGPU_OK;
}

int vsgpu_distances(vsgpu_store *s, const void *query, const uint32_t *ids, size_t n, double *out_scores) {
    if (!query || !ids || !out_scores) {
        set_error("vsgpu_distances: bad arguments");
        return VSGPU_ERR_ARG;
    }
    if (n == 0) return VSGPU_OK;
    for (size_t i = 0; i < n; i++)
        if (ids[i] >= s->count) {
            set_error("vsgpu_distances: id out of range");
            return VSGPU_ERR_ARG;
        }
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(resolve_pending_topk(s));
    s->stats = vsgpu_stats{};
    const size_t ssz = score_size(s);
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    VS_TRY(ensure_pinned(s, al(s->blob_bytes)));
    memcpy(s->pinned, query, s->blob_bytes);
    VS_TRY(ensure_scratch(s, s->q_raw, al(s->blob_bytes)));
    uint8_t *raw_dev = (uint8_t *)s->q_raw.ptr;
    VS_CUDA(cudaMemcpyAsync(raw_dev, s->pinned, s->blob_bytes, cudaMemcpyHostToDevice, s->stream));
    const void *q = nullptr;
    size_t qs = 0;
    const float *qn = nullptr;
    VS_TRY(stage_queries_device(s, raw_dev, 1, s->blob_bytes, &q, &qs, &qn));
    VS_TRY(ensure_scratch(s, s->out_dev, al(n * 4) + al(n * ssz)));
    uint32_t *d_ids = (uint32_t *)s->out_dev.ptr;
    uint8_t *d_sc = (uint8_t *)s->out_dev.ptr + al(n * 4);
    VS_CUDA(cudaMemcpyAsync(d_ids, ids, n * 4, cudaMemcpyHostToDevice, s->stream));
    VS_TRY(launch_exact_gather(s, q, 1, qs, qn, d_ids, n, nullptr, n, d_sc, n));
    if (ssz == 8) {
        VS_CUDA(cudaMemcpyAsync(out_scores, d_sc, n * 8, cudaMemcpyDeviceToHost, s->stream));
        VS_CUDA(cudaStreamSynchronize(s->stream));
    } else {
        std::vector<float> tmp(n);
        VS_CUDA(cudaMemcpyAsync(tmp.data(), d_sc, n * 4, cudaMemcpyDeviceToHost, s->stream));
        VS_CUDA(cudaStreamSynchronize(s->stream));
        for (size_t i = 0; i < n; i++) out_scores[i] = tmp[i];
    }
    return VSGPU_OK;
}

int vsgpu_last_stats(const vsgpu_store *s, vsgpu_stats *out) {
    if (!s || !out) return VSGPU_ERR_ARG;
    if (s->pend.active && s->ev1 && cudaEventQuery(s->ev1) == cudaSuccess) resolve_pending_topk(const_cast<vsgpu_store *>(s));
    *out = s->stats;
    // device-side timings become available once the call's last event has completed
    float ms = 0;
    if (out->total_ms == 0 && s->ev1 && cudaEventQuery(s->ev1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess)
        out->total_ms = ms;
    if (out->path == 0 && out->scan_ms == 0 && s->ev3 && cudaEventQuery(s->ev3) == cudaSuccess &&
        cudaEventElapsedTime(&ms, s->ev2, s->ev3) == cudaSuccess)
        out->scan_ms = ms;
    cudaGetLastError();
    return VSGPU_OK;
}

} // extern "C"
