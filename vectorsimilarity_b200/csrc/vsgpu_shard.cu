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

using namespace vsgpu;

extern "C" {

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
