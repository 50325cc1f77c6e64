// Sharded flat index, device side: what travels between the GPUs of one box and how per-shard top-k lists
// become one reply. The reference has no counterpart — its flat index is one in-process scan
// (algorithms/brute_force/brute_force.h:242-291); the merged order is the reference's reply order, ascending (score, label).
//
//   * a per-shard result travels as ONE packed list of 16-byte hits (label, fp32 score, flag) instead of separate score and
//     label arrays: one collective (or one set of peer stores) per batch. The flag of a query's first hit carries that
//     shard's "candidate buffer overflowed" bit, so every rank learns — without a host round trip — whether any shard has to
//     redo the query on the exact path;
//   * merge_packed_kernel: k-way merge by rank counting (lists are short: parts * k <= a few thousand), ORs the flags.
#include "vsgpu_internal.cuh"
#include <algorithm>
#include <cstring>
#include <vector>

namespace vsgpu {

struct __align__(16) PackedHit {
    uint64_t label;
    float score;
    uint32_t flag;
};

__device__ __forceinline__ uint32_t hit_key(float v) {
    uint32_t u = __float_as_uint(v);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void pack_hits_kernel(const float *__restrict__ scores, const uint64_t *__restrict__ labels,
                                 const uint32_t *__restrict__ flags, size_t nq, size_t k, PackedHit *__restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq * k; i += (size_t)gridDim.x * blockDim.x) {
        const size_t q = i / k;
        PackedHit h;
        h.label = labels[i];
        h.score = scores[i];
        h.flag = flags ? flags[q] : 0u;
        out[i] = h;
    }
}

// One block per query over `parts` sorted lists of k hits ([part][query][k]). Each hit counts the hits that sort before it
// (binary search in every other list; ties between lists by label, then by part) and lands at that rank.
__global__ void merge_packed_kernel(size_t parts, size_t nq, size_t k, const PackedHit *__restrict__ in, float *__restrict__ out_scores,
                                    uint64_t *__restrict__ out_labels, uint32_t *__restrict__ out_flags, uint32_t *__restrict__ any_flag) {
    const size_t q = blockIdx.x;
    const size_t total = parts * k;
    __shared__ uint32_t s_flag;
    if (threadIdx.x == 0) s_flag = 0;
    for (size_t i = threadIdx.x; i < k; i += blockDim.x) {
        out_scores[q * k + i] = __uint_as_float(0x7fc00000u);
        out_labels[q * k + i] = ~0ull;
    }
    __syncthreads();
    for (size_t e = threadIdx.x; e < total; e += blockDim.x) {
        const size_t p = e / k, j = e % k;
        const PackedHit h = in[(p * nq + q) * k + j];
        if (j == 0 && h.flag) atomicOr(&s_flag, 1u);
        if (h.label == ~0ull) continue;
        const uint32_t key = hit_key(h.score);
        size_t rank = 0;
        for (size_t p2 = 0; p2 < parts; p2++) {
            const PackedHit *l2 = in + (p2 * nq + q) * k;
            size_t lo = 0, hi = k;
            while (lo < hi) {
                const size_t mid = (lo + hi) / 2;
                const PackedHit m = l2[mid];
                bool before;
                if (m.label == ~0ull) before = false;
                else {
                    const uint32_t k2 = hit_key(m.score);
                    before = k2 < key || (k2 == key && (m.label < h.label || (m.label == h.label && p2 < p)));
                }
                if (before) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_scores[q * k + rank] = h.score;
            out_labels[q * k + rank] = h.label;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (out_flags) out_flags[q] = s_flag;
        if (s_flag && any_flag) atomicOr(any_flag, 1u);
    }
}

} // namespace vsgpu

// ---- one process, several devices (SURVEY.md §5 "Distributed communication backend", §8b "Additions") -------------------
// A group binds the stores of one sharded index. A batched query is enqueued shard by shard (each from its own host
// thread): queries H2D -> local top-k -> pack -> one peer copy of the packed list into the root device's gather buffer,
// all on the shard's stream; the root stream waits on one event per shard, merges and copies the reply to pinned memory.
// The only host wait is the one at the end of the batch.
struct vsgpu_group {
    std::vector<vsgpu_store *> stores;
    int root_dev = 0;
    cudaStream_t root_stream = nullptr;
    cudaEvent_t root_ev0 = nullptr, root_ev1 = nullptr;
    struct Shard {
        vsgpu::Scratch q, sc, lab, packed;
        cudaEvent_t done = nullptr;
    };
    std::vector<Shard> shards;
    vsgpu::Scratch gathered, out_sc, out_lab, flag; // root device
    void *pin_q = nullptr;                          // pinned: processed queries (written once, read by every shard's H2D)
    size_t pin_q_bytes = 0;
    void *pin_out = nullptr;                        // pinned: [labels | scores | any flag]
    size_t pin_out_bytes = 0;
    float last_ms = 0.f;
};

namespace vsgpu {
static int ensure_dev(int device, cudaStream_t st, Scratch &sc, size_t bytes) {
    if (bytes <= sc.bytes) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(device));
    if (sc.ptr) {
        VS_CUDA(cudaStreamSynchronize(st));
        VS_CUDA(cudaFree(sc.ptr));
        sc.ptr = nullptr;
        sc.bytes = 0;
    }
    const size_t want = (std::max<size_t>(bytes, 4096) + 255) / 256 * 256;
    VS_CUDA(cudaMalloc(&sc.ptr, want));
    sc.bytes = want;
    return VSGPU_OK;
}
static int ensure_pin(void *&p, size_t &have, size_t bytes) {
    if (bytes <= have) return VSGPU_OK;
    if (p) VS_CUDA(cudaFreeHost(p));
    p = nullptr;
    have = 0;
    const size_t want = std::max<size_t>(bytes, 65536);
    VS_CUDA(cudaMallocHost(&p, want));
    have = want;
    return VSGPU_OK;
}
} // namespace vsgpu

using namespace vsgpu;

extern "C" {

vsgpu_group *vsgpu_group_create(vsgpu_store **stores, size_t n) {
    if (!stores || n == 0) {
        set_error("vsgpu_group_create: no stores");
        return nullptr;
    }
    for (size_t i = 0; i < n; i++)
        if (!stores[i] || stores[i]->type == VSGPU_FLOAT64 || stores[i]->type != stores[0]->type || stores[i]->dim != stores[0]->dim ||
            stores[i]->metric != stores[0]->metric) {
            set_error("vsgpu_group_create: stores must share type / dim / metric (fp64 indexes are not sharded)");
            return nullptr;
        }
    auto *g = new vsgpu_group();
    g->stores.assign(stores, stores + n);
    g->shards.resize(n);
    g->root_dev = stores[0]->device;
    bool ok = cudaSetDevice(g->root_dev) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&g->root_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreate(&g->root_ev0) == cudaSuccess && cudaEventCreate(&g->root_ev1) == cudaSuccess;
    for (size_t i = 0; ok && i < n; i++) {
        ok = cudaSetDevice(stores[i]->device) == cudaSuccess;
        ok = ok && cudaEventCreateWithFlags(&g->shards[i].done, cudaEventDisableTiming) == cudaSuccess;
        // direct NVLink stores / loads between every pair of shard devices (peer copies fall back to staging otherwise)
        for (size_t j = 0; ok && j < n; j++) {
            if (stores[j]->device == stores[i]->device) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, stores[i]->device, stores[j]->device);
            if (can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(stores[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = false;
                cudaGetLastError();
            }
        }
    }
    if (!ok) {
        set_error("vsgpu_group_create: CUDA resource creation failed");
        vsgpu_group_destroy(g);
        return nullptr;
    }
    return g;
}

void vsgpu_group_destroy(vsgpu_group *g) {
    if (!g) return;
    for (size_t i = 0; i < g->shards.size(); i++) {
        cudaSetDevice(g->stores[i]->device);
        cudaStreamSynchronize(g->stores[i]->stream);
        for (Scratch *sc : {&g->shards[i].q, &g->shards[i].sc, &g->shards[i].lab, &g->shards[i].packed})
            if (sc->ptr) cudaFree(sc->ptr);
        if (g->shards[i].done) cudaEventDestroy(g->shards[i].done);
    }
    cudaSetDevice(g->root_dev);
    if (g->root_stream) cudaStreamSynchronize(g->root_stream);
    for (Scratch *sc : {&g->gathered, &g->out_sc, &g->out_lab, &g->flag})
        if (sc->ptr) cudaFree(sc->ptr);
    if (g->pin_q) cudaFreeHost(g->pin_q);
    if (g->pin_out) cudaFreeHost(g->pin_out);
    if (g->root_ev0) cudaEventDestroy(g->root_ev0);
    if (g->root_ev1) cudaEventDestroy(g->root_ev1);
    if (g->root_stream) cudaStreamDestroy(g->root_stream);
    delete g;
}

size_t vsgpu_group_size(const vsgpu_group *g) { return g->stores.size(); }

// Step 1 (one thread): stage nq processed query blobs (HOST, qstride apart) in the group's pinned buffer and size the
// root-side buffers.
int vsgpu_group_topk_begin(vsgpu_group *g, const void *queries, size_t nq, size_t qstride, size_t k) {
    vsgpu_store *s0 = g->stores[0];
    if (!queries || nq == 0 || k == 0 || qstride < s0->blob_bytes) {
        set_error("vsgpu_group_topk_begin: bad arguments");
        return VSGPU_ERR_ARG;
    }
    VS_CUDA(cudaSetDevice(g->root_dev));
    VS_CUDA(cudaStreamSynchronize(g->root_stream)); // the previous batch has left the pinned buffers
    VS_TRY(ensure_pin(g->pin_q, g->pin_q_bytes, nq * s0->blob_bytes));
    for (size_t q = 0; q < nq; q++) memcpy((uint8_t *)g->pin_q + q * s0->blob_bytes, (const uint8_t *)queries + q * qstride, s0->blob_bytes);
    const size_t parts = g->stores.size();
    VS_TRY(ensure_dev(g->root_dev, g->root_stream, g->gathered, parts * nq * k * sizeof(PackedHit)));
    VS_TRY(ensure_dev(g->root_dev, g->root_stream, g->out_sc, nq * k * 4));
    VS_TRY(ensure_dev(g->root_dev, g->root_stream, g->out_lab, nq * k * 8));
    VS_TRY(ensure_dev(g->root_dev, g->root_stream, g->flag, 256));
    VS_TRY(ensure_pin(g->pin_out, g->pin_out_bytes, nq * k * 12 + 256));
    VS_CUDA(cudaEventRecord(g->root_ev0, g->root_stream));
    return VSGPU_OK;
}

// Step 2 (one call per shard, from any thread — calls on different shards run concurrently): enqueue shard i's scan and
// the push of its packed top-k into slot i of the root's gather buffer. `repush`: the scan already ran (overflowed
// queries were just redone by vsgpu_store_sync); only pack and push again.
int vsgpu_group_topk_shard(vsgpu_group *g, size_t i, size_t nq, size_t k, unsigned flags, int repush) {
    if (i >= g->stores.size()) return VSGPU_ERR_ARG;
    vsgpu_store *s = g->stores[i];
    vsgpu_group::Shard &sh = g->shards[i];
    VS_CUDA(cudaSetDevice(s->device));
    VS_TRY(ensure_dev(s->device, s->stream, sh.q, nq * s->blob_bytes));
    VS_TRY(ensure_dev(s->device, s->stream, sh.sc, nq * k * 4));
    VS_TRY(ensure_dev(s->device, s->stream, sh.lab, nq * k * 8));
    VS_TRY(ensure_dev(s->device, s->stream, sh.packed, nq * k * sizeof(PackedHit)));
    if (!repush) {
        VS_CUDA(cudaMemcpyAsync(sh.q.ptr, g->pin_q, nq * s->blob_bytes, cudaMemcpyHostToDevice, s->stream));
        VS_TRY(vsgpu_topk_device(s, sh.q.ptr, nq, s->blob_bytes, k, flags, (uint64_t *)sh.lab.ptr, sh.sc.ptr, nullptr));
    }
    VS_TRY(vsgpu_pack_topk_device(s, nq, k, (const float *)sh.sc.ptr, (const uint64_t *)sh.lab.ptr, sh.packed.ptr));
    uint8_t *slot = (uint8_t *)g->gathered.ptr + i * nq * k * sizeof(PackedHit);
    if (s->device == g->root_dev)
        VS_CUDA(cudaMemcpyAsync(slot, sh.packed.ptr, nq * k * sizeof(PackedHit), cudaMemcpyDeviceToDevice, s->stream));
    else
        VS_CUDA(cudaMemcpyPeerAsync(slot, g->root_dev, sh.packed.ptr, s->device, nq * k * sizeof(PackedHit), s->stream));
    VS_CUDA(cudaEventRecord(sh.done, s->stream));
    return VSGPU_OK;
}

// Step 3 (one thread): the root stream waits for every shard's push, merges and copies the reply to the host. Returns 1
// (and leaves the outputs untouched) when some shard flagged an overflowed query: the caller then runs
// vsgpu_store_sync on every shard, step 2 with repush = 1 and this step again.
int vsgpu_group_topk_finish(vsgpu_group *g, size_t nq, size_t k, uint64_t *out_labels, double *out_scores) {
    VS_CUDA(cudaSetDevice(g->root_dev));
    for (auto &sh : g->shards) VS_CUDA(cudaStreamWaitEvent(g->root_stream, sh.done, 0));
    VS_CUDA(cudaMemsetAsync(g->flag.ptr, 0, 4, g->root_stream));
    VS_TRY(vsgpu_merge_packed_device(g->root_dev, g->root_stream, g->stores.size(), nq, k, g->gathered.ptr, (float *)g->out_sc.ptr,
                                     (uint64_t *)g->out_lab.ptr, nullptr, (uint32_t *)g->flag.ptr));
    uint8_t *pin = (uint8_t *)g->pin_out;
    const size_t o_sc = nq * k * 8, o_fl = o_sc + nq * k * 4;
    VS_CUDA(cudaMemcpyAsync(pin, g->out_lab.ptr, nq * k * 8, cudaMemcpyDeviceToHost, g->root_stream));
    VS_CUDA(cudaMemcpyAsync(pin + o_sc, g->out_sc.ptr, nq * k * 4, cudaMemcpyDeviceToHost, g->root_stream));
    VS_CUDA(cudaMemcpyAsync(pin + o_fl, g->flag.ptr, 4, cudaMemcpyDeviceToHost, g->root_stream));
    VS_CUDA(cudaEventRecord(g->root_ev1, g->root_stream));
    VS_CUDA(cudaStreamSynchronize(g->root_stream));
    cudaEventElapsedTime(&g->last_ms, g->root_ev0, g->root_ev1);
    if (*(const uint32_t *)(pin + o_fl)) return 1;
    if (out_labels) memcpy(out_labels, pin, nq * k * 8);
    if (out_scores) {
        const float *sc = (const float *)(pin + o_sc);
        for (size_t i = 0; i < nq * k; i++) out_scores[i] = (double)sc[i];
    }
    return VSGPU_OK;
}

float vsgpu_group_last_ms(const vsgpu_group *g) { return g->last_ms; }

// device that owns a device pointer, or -1
int vsgpu_pointer_device(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess || a.type != cudaMemoryTypeDevice) {
        cudaGetLastError();
        return -1;
    }
    return a.device;
}

size_t vsgpu_packed_hit_bytes(void) { return sizeof(PackedHit); }

// Pack the [nq][k] (score fp32, label) lists a vsgpu_topk_device call on this store produced, together with the call's
// per-query overflow flags, into `out` ([nq][k] PackedHit, DEVICE). Enqueued on the store's stream.
int vsgpu_pack_topk_device(vsgpu_store *s, size_t nq, size_t k, const float *scores, const uint64_t *labels, void *out) {
    if (!s || !scores || !labels || !out) {
        set_error("vsgpu_pack_topk_device: bad arguments");
        return VSGPU_ERR_ARG;
    }
    if (nq == 0 || k == 0) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(s->device));
    const uint32_t *flags = (s->pend.active && s->pend.nq == nq) ? (const uint32_t *)s->ovf.ptr : nullptr;
    const unsigned blocks = (unsigned)std::min<size_t>((nq * k + 255) / 256, 2048);
    pack_hits_kernel<<<blocks, 256, 0, s->stream>>>(scores, labels, flags, nq, k, (PackedHit *)out);
    VS_CUDA(cudaGetLastError());
    s->stats.kernel_launches++;
    return VSGPU_OK;
}

// Merge `parts` packed lists ([parts][nq][k], DEVICE) into [nq][k] scores / labels by ascending (score, label).
// out_flags ([nq], may be NULL): OR over the parts of each query's overflow flag; any_flag (one u32, may be NULL, must
// be zero on entry): OR over all queries.
int vsgpu_merge_packed_device(int device, void *stream, size_t parts, size_t nq, size_t k, const void *packed, float *out_scores,
                              uint64_t *out_labels, uint32_t *out_flags, uint32_t *any_flag) {
    if (parts == 0 || nq == 0 || k == 0) return VSGPU_OK;
    VS_CUDA(cudaSetDevice(device));
    const unsigned threads = (unsigned)std::min<size_t>(1024, std::max<size_t>(32, (parts * k + 31) / 32 * 32));
    merge_packed_kernel<<<(unsigned)nq, threads, 0, (cudaStream_t)stream>>>(parts, nq, k, (const PackedHit *)packed, out_scores,
                                                                           out_labels, out_flags, any_flag);
    VS_CUDA(cudaGetLastError());
    return VSGPU_OK;
}

} // extern "C"
