/* include/vsgpu.h — the thin C-ABI between the host index code and the sm_100a CUDA kernels
 * (libvsgpu.so). Plain pointers and sizes only; no C++/torch types cross this line.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference/src/VecSim):
 *   vsgpu_store_*      containers/data_blocks_container.{h,cpp} + data_block.{h,cpp}: the
 *                      "VectorBlock" store (addElement :30-39, updateElement, removeElement :41-56)
 *                      and the idToLabelMapping of algorithms/brute_force/brute_force.h:35,174-224.
 *   vsgpu_topk         the scan + max-heap loop of BruteForceIndex::topKQuery
 *                      (algorithms/brute_force/brute_force.h:262-288), for a batch of queries.
 *   vsgpu_range        BruteForceIndex::rangeQuery's scan (brute_force.h:304-321).
 *   vsgpu_scores       BFS_BatchIterator::calculateScores (bfs_batch_iterator.h:24-41).
 *   vsgpu_distances    calcDistance over chosen ids (vec_sim_index.h:175-190) — ad-hoc BF and
 *                      getDistanceFrom_Unsafe (brute_force_single.h:200-212).
 *   vsgpu_hnsw_*       HNSWIndex::topKQuery's traversal (algorithms/hnsw/hnsw.h:530-613,
 *                      1210-1258, 1967-2084) over a graph built on the host.
 * Distances reproduce the reference's x86 AVX512 dispatch tier bit for bit.
 *
 * All functions return VSGPU_OK (0) or a negative error; vsgpu_last_error() gives the text.
 * A store is bound to one device and one stream; calls on one store must be serialised by the
 * caller (the host index holds a mutex), different stores are independent.
 */
#ifndef VSGPU_H
#define VSGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* numbering = VecSimType / VecSimMetric (vec_sim_common.h:60-69,84-88) */
enum { VSGPU_FLOAT32 = 0, VSGPU_FLOAT64 = 1, VSGPU_BFLOAT16 = 2, VSGPU_FLOAT16 = 3, VSGPU_INT8 = 4, VSGPU_UINT8 = 5 };
enum { VSGPU_L2 = 0, VSGPU_IP = 1, VSGPU_COSINE = 2 };

enum {
    VSGPU_OK = 0,
    VSGPU_ERR_CUDA = -1,     /* a CUDA call failed */
    VSGPU_ERR_ARG = -2,      /* bad argument */
    VSGPU_ERR_NOMEM = -3,    /* device allocation failed */
    VSGPU_ERR_NODEVICE = -4, /* no usable sm_100 device */
    VSGPU_ERR_OVERFLOW = -5  /* caller's output buffer too small (range query) */
};

/* vsgpu_topk flags */
enum {
    VSGPU_TOPK_AUTO = 0,       /* tensor-core coarse pass + exact re-rank when the shape allows */
    VSGPU_TOPK_EXACT_ONLY = 1, /* force the exact SIMT scan */
    VSGPU_TOPK_TENSOR_ONLY = 2 /* fail instead of falling back (tests / benchmarks) */
};

typedef struct vsgpu_store vsgpu_store;

const char *vsgpu_last_error(void); /* per calling thread */
void vsgpu_set_last_error(const char *msg); /* hand an error text from a worker thread to the thread that reports it */
int vsgpu_device_count(void);
/* free / total bytes of HBM on `device` */
int vsgpu_mem_info(int device, size_t *free_bytes, size_t *total_bytes);

/* A store holds processed rows (what the reference keeps in its DataBlocks): dim elements of
 * `type`, plus — for int8/uint8 cosine — the fp32 norm that the reference appends to the blob.
 * `capacity_hint` rows are reserved up front (0 = grow on demand). */
vsgpu_store *vsgpu_store_create(int device, int type, int metric, size_t dim, size_t capacity_hint);
void vsgpu_store_destroy(vsgpu_store *s);
size_t vsgpu_store_size(const vsgpu_store *s);
size_t vsgpu_store_row_bytes(const vsgpu_store *s); /* = VecSimParams_GetStoredDataSize */
size_t vsgpu_store_device_bytes(const vsgpu_store *s);

/* Append n processed rows from HOST memory (`stride` bytes apart) with their labels. */
int vsgpu_store_append(vsgpu_store *s, const void *rows, size_t stride, const uint64_t *labels, size_t n);
/* Same, rows and labels already in DEVICE memory of the store's device (bulk loaders).
 * labels == NULL: row i gets first_label + i. For int8/uint8 cosine `norms` (device, fp32) replaces
 * the appended norm; NULL = compute on device. */
int vsgpu_store_append_device(vsgpu_store *s, const void *rows, size_t stride, const uint64_t *labels,
                              uint64_t first_label, const float *norms, size_t n);
/* Overwrite row `id` (label update in place). */
int vsgpu_store_update(vsgpu_store *s, size_t id, const void *row, uint64_t label);
/* Delete-by-swap (brute_force.h:195-224): row `src` (the last one) moves to `dst`; count -= 1. */
int vsgpu_store_remove_swap(vsgpu_store *s, size_t dst);
/* Drop the rows [new_count, size()) (roll back an append whose follow-up step failed). */
int vsgpu_store_truncate(vsgpu_store *s, size_t new_count);
/* Copy rows back to the host (tests, serialisation). */
int vsgpu_store_read(const vsgpu_store *s, size_t first, size_t n, void *rows, size_t stride, uint64_t *labels);

/* Batched top-k. `queries`: nq processed query blobs in HOST memory, `qstride` bytes apart.
 * Outputs (HOST, any may be NULL): [nq][k] row-major, padded with label=UINT64_MAX/score=NaN/
 * id=UINT32_MAX; out_counts[q] = number of valid entries = min(k, size).
 * Order: ascending (score, internal id). With labels that grow with the internal id this IS the
 * reference's result order; the host index resolves the general tie rule from out_ids. */
int vsgpu_topk(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t k, unsigned flags,
               uint64_t *out_labels, double *out_scores, uint32_t *out_ids, uint32_t *out_counts);
/* Same with queries and outputs in DEVICE memory (scores as the index's DistType widened to
 * double is a host concern: here fp32, or fp64 for FLOAT64 stores, in `out_scores`). Work is
 * enqueued on the store's stream; call vsgpu_store_sync before reading. */
int vsgpu_topk_device(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t k,
                      unsigned flags, uint64_t *out_labels, void *out_scores, uint32_t *out_ids);
/* Phased variant for sharded callers (`world` shards, one store each). A tensor-path call runs its coarse pass in `rounds`
 * phases — _begin the first, every _next one more — and after each writes this shard's bounds to `bounds` ([2 nq] fp32,
 * DEVICE): bounds[q] bounds its k-th best score from below, -bounds[nq + q] its ceil(k / world)-th best. Between the calls
 * the caller reduces the whole buffer with MAX over the shards (one small all-reduce); max(bounds[q], -bounds[nq + q]) is then
 * a lower bound of the k-th best score overall (world * ceil(k / world) >= k rows reach the smallest of the shards' second
 * values). Every shard admits against it in its next phase and prunes against it in _finish, before the exact re-rank.
 * Sequence: _begin, (reduce, _next) x (rounds - 1), reduce, _finish. `rounds` = vsgpu_topk_rounds() of a row count all the
 * shards agree on — every shard runs exactly that many phases whatever its own size. Calls that take another path do all
 * their work in _begin and leave neutral bounds. Same outputs as vsgpu_topk_device. world < 2 or bounds == NULL: plain call. */
int vsgpu_topk_device_begin(vsgpu_store *s, const void *queries, size_t nq, size_t qstride, size_t k, unsigned flags,
                            uint64_t *out_labels, void *out_scores, uint32_t *out_ids, unsigned world, unsigned rounds,
                            float *bounds);
int vsgpu_topk_device_next(vsgpu_store *s, float *bounds);
int vsgpu_topk_device_finish(vsgpu_store *s, const float *bounds);
size_t vsgpu_topk_rounds(size_t rows_per_shard, size_t k, unsigned world);
int vsgpu_store_sync(vsgpu_store *s);
void *vsgpu_store_stream(vsgpu_store *s); /* cudaStream_t */
/* Run the store's work on a stream of the caller's (a framework's pooled stream, say) from now on; the store never
 * destroys it, and it must outlive the store. */
int vsgpu_store_set_stream(vsgpu_store *s, void *stream);

/* Range scan for ONE query: every row with score <= radius (already cast to the DistType by the
 * caller). Unordered. Returns VSGPU_ERR_OVERFLOW and sets *out_count to the needed capacity when
 * `cap` is too small. */
int vsgpu_range(vsgpu_store *s, const void *query, double radius, size_t cap, uint64_t *out_labels,
                double *out_scores, uint32_t *out_ids, size_t *out_count);
/* All scores of ONE query, by internal id (out_scores: size() doubles, HOST). */
int vsgpu_scores(vsgpu_store *s, const void *query, double *out_scores);
/* Scores of chosen rows (ids: HOST). */
int vsgpu_distances(vsgpu_store *s, const void *query, const uint32_t *ids, size_t n, double *out_scores);

/* Counters of the last vsgpu_topk call on this store (for benchmarks and tests). */
typedef struct {
    uint32_t path;            /* 0 exact scan, 1 tensor coarse + re-rank */
    uint32_t kernel_launches; /* kernels launched by the call */
    uint64_t candidates;      /* rows re-ranked exactly (tensor path) */
    uint32_t fallback_queries;/* queries redone on the exact path (candidate overflow) */
    float scan_ms;            /* device time of the dominant scan kernel(s), CUDA events */
    float total_ms;           /* device time of the whole call */
} vsgpu_stats;
int vsgpu_last_stats(const vsgpu_store *s, vsgpu_stats *out);

/* Merge per-shard top-k lists (DEVICE memory): `parts` lists of [nq][k] (score fp32/fp64, label)
 * laid out part-major, into [nq][k] by ascending (score, label). Used after the all-gather of the
 * sharded flat index. dtype_f64 != 0 for FLOAT64 indexes. */
int vsgpu_merge_topk_device(int device, void *stream, int dtype_f64, size_t parts, size_t nq, size_t k,
                            const void *scores, const uint64_t *labels, void *out_scores,
                            uint64_t *out_labels);

/* Sharded flat index: a shard's result travels as ONE list of 16-byte hits {u64 label; f32 score; u32 flag}
 * — one collective per batch. The flag of a query's first hit carries "this shard's candidate buffer overflowed" (tensor
 * path), so every rank learns without a host round trip whether the query has to be redone on the exact path. */
size_t vsgpu_packed_hit_bytes(void);
/* Pack the lists the last vsgpu_topk_device call on `s` wrote ([nq][k] fp32 scores / labels, DEVICE) into `out`
 * ([nq][k] hits, DEVICE), on the store's stream. */
int vsgpu_pack_topk_device(vsgpu_store *s, size_t nq, size_t k, const float *scores, const uint64_t *labels, void *out);
/* Merge `parts` packed lists ([parts][nq][k], DEVICE) by ascending (score, label). out_flags [nq] (may be NULL): OR of the
 * parts' overflow flags per query; any_flag (one u32, zero on entry, may be NULL): OR over the queries. */
int vsgpu_merge_packed_device(int device, void *stream, size_t parts, size_t nq, size_t k, const void *packed,
                              float *out_scores, uint64_t *out_labels, uint32_t *out_flags, uint32_t *any_flag);

/* ---- one process, several devices ------------------------------------------------------------------------------------
 * A group binds the stores (one per device) of a sharded flat index. A batched top-k runs as:
 *   vsgpu_group_topk_begin   (one thread)   stage the processed queries in pinned memory
 *   vsgpu_group_topk_shard   (per shard, concurrently from several threads)  H2D, local top-k, pack, peer copy of the packed
 *                            list into the root device's gather buffer — all on the shard's stream
 *   vsgpu_group_topk_finish  (one thread)   root stream waits for the shards' events, merges, copies the reply to the host;
 *                            returns 1 when a shard flagged an overflowed query (then: vsgpu_store_sync on every shard,
 *                            _shard with repush = 1, _finish again)
 * The first store's device is the root. */
typedef struct vsgpu_group vsgpu_group;
vsgpu_group *vsgpu_group_create(vsgpu_store **stores, size_t n);
void vsgpu_group_destroy(vsgpu_group *g);
size_t vsgpu_group_size(const vsgpu_group *g);
int vsgpu_group_topk_begin(vsgpu_group *g, const void *queries, size_t nq, size_t qstride, size_t k);
int vsgpu_group_topk_shard(vsgpu_group *g, size_t shard, size_t nq, size_t k, unsigned flags, int repush);
int vsgpu_group_topk_finish(vsgpu_group *g, size_t nq, size_t k, uint64_t *out_labels, double *out_scores);
float vsgpu_group_last_ms(const vsgpu_group *g); /* device time of the last batch, root stream */
int vsgpu_pointer_device(const void *device_ptr); /* owning device of a device pointer, -1 if not device memory */

/* ---- HNSW (algorithms/hnsw/hnsw.h) -----------------------------------------------------------
 * A graph over the rows of a store (internal id = row index). Level-0 records hold up to 2M links,
 * upper levels M. Traversal results are identical to HNSWIndex::topKQuery / rangeQuery on the same
 * graph (same admission order, same heaps' total order, bit-exact distances); the builder inserts
 * sequentially in id order and reproduces the reference's single-threaded graph. */
typedef struct vsgpu_hnsw vsgpu_hnsw;
/* M in [2,256]; ef_construction is raised to M like hnsw.h:1632-1633. */
vsgpu_hnsw *vsgpu_hnsw_create(vsgpu_store *s, size_t M, size_t ef_construction);
void vsgpu_hnsw_destroy(vsgpu_hnsw *g);
size_t vsgpu_hnsw_size(const vsgpu_hnsw *g);
/* Several rows per label (HNSWIndex_Multi, hnsw_multi.h:62-71,103-106): top-k and the batch iterator key their result set by
 * label (updatable_max_heap, utils/updatable_heap.h:66-111) — each label once, with its best score. */
void vsgpu_hnsw_set_multi(vsgpu_hnsw *g, int multi);
size_t vsgpu_hnsw_device_bytes(const vsgpu_hnsw *g);
int vsgpu_hnsw_entry(const vsgpu_hnsw *g, long *entry, long *max_level); /* -1/-1 when empty */
/* Index the next n rows of the store (ids size()..size()+n-1, appended beforehand) with the given
 * top levels (HOST; drawn by the caller as hnsw.h:418-422 does). insertElementToGraph :1567-1602. */
int vsgpu_hnsw_insert(vsgpu_hnsw *g, size_t n, const uint32_t *levels);
/* Bulk load / read back a graph (HOST buffers). l0: n x (2M+1) u32 = count then links; upper: one
 * (M+1)-u32 record per (node, level>=1), nodes in id order, levels ascending. */
int vsgpu_hnsw_import(vsgpu_hnsw *g, size_t n, const uint32_t *levels, const uint32_t *l0, const uint32_t *upper,
                      size_t upper_records, long entry, long max_level);
int vsgpu_hnsw_export(const vsgpu_hnsw *g, uint32_t *levels, uint32_t *l0, uint32_t *upper, size_t upper_cap_records,
                      size_t *upper_records);
/* One node's records (HOST): level 0 record (2M+1 words) then one (M+1)-word record per upper level.
 * records == NULL: only the level. VSGPU_ERR_OVERFLOW when cap_words is too small. */
int vsgpu_hnsw_node(const vsgpu_hnsw *g, size_t id, uint32_t *level_out, uint32_t *records, size_t cap_words);
/* markDelete / unmark (VecSimIndexTombstone): deleted nodes are traversed but never returned. */
int vsgpu_hnsw_set_deleted(vsgpu_hnsw *g, size_t id, int deleted);
/* Batched top-k, ef = max(ef, k) (hnsw.h:2072). Layout and padding as vsgpu_topk; results ascending
 * (score, label). out_counts[q] <= k. */
int vsgpu_hnsw_topk(vsgpu_hnsw *g, const void *queries, size_t nq, size_t qstride, size_t k, size_t ef,
                    uint64_t *out_labels, double *out_scores, uint32_t *out_ids, uint32_t *out_counts);
/* DEVICE queries/outputs (scores in the DistType); synchronises the store's stream before returning. */
int vsgpu_hnsw_topk_device(vsgpu_hnsw *g, const void *queries, size_t nq, size_t qstride, size_t k, size_t ef,
                           uint64_t *out_labels, void *out_scores, uint32_t *out_ids, uint32_t *out_counts);
/* Range search for ONE query (hnsw.h:2086-2200). Unordered; VSGPU_ERR_OVERFLOW + needed count as vsgpu_range. */
int vsgpu_hnsw_range(vsgpu_hnsw *g, const void *query, double radius, double epsilon, size_t cap, uint64_t *out_labels,
                     double *out_scores, uint32_t *out_ids, size_t *out_count);
/* Resumable batch iterator (hnsw_batch_iterator.h:59-267): keeps the traversal state (visited set, candidate and
 * spare-result heaps, lower bound) on the device between calls; every `next` returns what
 * HNSW_BatchIterator::getNextResults(n) returns, ascending (score, label). `ef` = the query's efRuntime.
 * `label_count` = indexLabelCount() at call time (depletion rule :243-245). */
typedef struct vsgpu_hnsw_iter vsgpu_hnsw_iter;
vsgpu_hnsw_iter *vsgpu_hnsw_iter_create(vsgpu_hnsw *g, const void *query, size_t ef);
void vsgpu_hnsw_iter_destroy(vsgpu_hnsw_iter *it);
int vsgpu_hnsw_iter_reset(vsgpu_hnsw_iter *it);
/* Multi-value graphs (HNSWMulti_BatchIterator::returned, hnsw_multi_batch_iterator.h:39-57): after a batch, the caller names
 * every row of every label that batch returned; later batches skip them. HOST ids. */
int vsgpu_hnsw_iter_mark_returned(vsgpu_hnsw_iter *it, const uint32_t *ids, size_t n);
int vsgpu_hnsw_iter_next(vsgpu_hnsw_iter *it, size_t n_res, size_t label_count, uint64_t *out_labels, double *out_scores,
                         uint32_t *out_ids, size_t *out_count, int *depleted);
/* Counters of the last traversal / insert call: distance evaluations, expanded nodes, device ms. */
int vsgpu_hnsw_last_stats(const vsgpu_hnsw *g, unsigned long long *dist_evals, unsigned long long *hops, float *ms);

#ifdef __cplusplus
}
#endif
#endif
