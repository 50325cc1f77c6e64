// Host-side index objects behind the VecSimIndex_* C API (include/vecsim_b200.h).
// Mirrors the reference's VecSimIndexInterface / BruteForceIndex_Single split
// (/root/reference/src/VecSim/vec_sim_interface.h:23-243, algorithms/brute_force/brute_force.h,
// brute_force_single.h) but keeps only host bookkeeping here: rows live in HBM behind vsgpu_store.
#pragma once
#include "../../../include/vecsim_b200.h"
#include "../../../include/vsgpu.h"
#include <cstdint>
#include <functional>
#include <memory>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <shared_mutex>
#include <random>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace vsb {
// Host allocations that live as long as an index, a reply or an iterator go through the caller's hooks
// (VecSim_SetMemoryFunctions, /root/reference/src/VecSim/vec_sim.h:281, memory/vecsim_malloc.cpp:27-65): RediSearch hands
// in RedisModule_Alloc / Free so the memory shows up in Redis' accounting. A 16-byte header keeps the size and the free
// function that matches the allocation (hooks may be installed after some objects exist); the live total is what
// hook_bytes() reports.
void *hook_alloc(size_t n);
void hook_free(void *p) noexcept;
size_t hook_bytes();
template <class T> struct HookAlloc {
    using value_type = T;
    HookAlloc() = default;
    template <class U> HookAlloc(const HookAlloc<U> &) {}
    T *allocate(size_t n) { return static_cast<T *>(hook_alloc(n * sizeof(T))); }
    void deallocate(T *p, size_t) noexcept { hook_free(p); }
    template <class U> bool operator==(const HookAlloc<U> &) const { return true; }
    template <class U> bool operator!=(const HookAlloc<U> &) const { return false; }
};
template <class T> using hvec = std::vector<T, HookAlloc<T>>;
template <class K, class V> using hmap = std::unordered_map<K, V, std::hash<K>, std::equal_to<K>, HookAlloc<std::pair<const K, V>>>;
struct Hooked { // objects handed to the caller (indexes, replies, iterators)
    static void *operator new(size_t n) { return hook_alloc(n); }
    static void operator delete(void *p) noexcept { hook_free(p); }
};
} // namespace vsb

struct VecSimQueryResult {
    size_t id;
    double score;
};
struct VecSimQueryReply : vsb::Hooked {
    std::vector<VecSimQueryResult> results;
    VecSimQueryReply_Code code = VecSim_QueryReply_OK;
};
struct VecSimQueryReply_Iterator {
    VecSimQueryReply *reply;
    size_t pos;
};

namespace vsb {
struct Globals {
    timeoutCallbackFunction timeout_cb = nullptr;
    logCallbackFunction log_cb = nullptr;
    VecSimMemoryFunctions mem{};
    bool mem_set = false;
    VecSimWriteMode write_mode = VecSim_WriteAsync;
    int device = 0;
    std::vector<int> devices; // VecSimGPU_Configure: more than one entry -> new flat indexes are sharded over them
    int topk_mode = 0;
};
Globals &globals();
inline bool timed_out(void *ctx) { return globals().timeout_cb && globals().timeout_cb(ctx) != 0; }
size_t type_size(VecSimType t);
size_t stored_size(VecSimType t, size_t dim, VecSimMetric m);
void normalize_blob(void *blob, size_t dim, VecSimType type); // VecSim_Normalize

} // namespace vsb

struct VecSimDebugInfoIterator {
    std::vector<VecSim_InfoField> fields;
    size_t pos = 0;
};

struct VecSimBatchIterator : vsb::Hooked {
    virtual ~VecSimBatchIterator() = default;
    virtual VecSimQueryReply *next(size_t n, VecSimQueryReply_Order order) = 0;
    virtual bool hasNext() = 0;
    virtual void reset() = 0;
};

struct VecSimIndexInterface : vsb::Hooked {
    virtual ~VecSimIndexInterface() = default;
    virtual int addVector(const void *blob, size_t label) = 0;
    virtual long addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) = 0;
    virtual int deleteVector(size_t label) = 0;
    virtual double getDistanceFrom(size_t label, const void *blob) = 0;
    virtual size_t indexSize() = 0;
    virtual size_t indexLabelCount() = 0;
    virtual VecSimQueryReply *topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) = 0;
    virtual int topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *labels,
                          double *scores, uint32_t *counts) = 0;
    virtual VecSimQueryReply *rangeQuery(const void *blob, double radius, VecSimQueryParams *qp,
                                         VecSimQueryReply_Order order) = 0;
    virtual VecSimBatchIterator *newBatchIterator(const void *blob, VecSimQueryParams *qp) = 0;
    virtual VecSimIndexBasicInfo basicInfo() = 0;
    virtual VecSimIndexDebugInfo debugInfo() = 0;
    virtual VecSimIndexStatsInfo statsInfo() = 0;
    virtual bool preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) = 0;
    virtual void setLastSearchMode(VecSearchMode m) = 0;
    virtual void exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) = 0;
    virtual std::vector<uint8_t> preprocessQuery(const void *blob) = 0;
    virtual vsgpu_store *deviceStore() = 0;
    virtual void lastStats(vsgpu_stats *out) = 0;
    // VecSimDebug_GetElementNeighborsInHNSWGraph (HNSW only)
    virtual int elementNeighbors(size_t, int ***out) {
        *out = nullptr;
        return VecSimDebugCommandCode_BadIndex;
    }
};

namespace vsb {

class FlatIndex final : public VecSimIndexInterface {
  public:
    FlatIndex(const BFParams &p, void *logCtx, int device = -1); // device < 0: the process-wide default (VecSimGPU_SetDevice)
    ~FlatIndex() override;
    bool ok() const { return store_ != nullptr; }
    // labels grow with the internal id: the device's (score, id) order is the reference's (score, label) reply order
    bool labelsMonotone() {
        std::lock_guard<std::mutex> g(mu_);
        return labels_monotone_;
    }

    int addVector(const void *blob, size_t label) override;
    long addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) override;
    int deleteVector(size_t label) override;
    double getDistanceFrom(size_t label, const void *blob) override;
    size_t indexSize() override { return count_; }
    size_t indexLabelCount() override { return count_; }
    bool hasLabel(size_t label) {
        std::lock_guard<std::mutex> g(mu_);
        idType id;
        return findId(label, &id);
    }
    long appendDeviceRows(const void *dev_rows, size_t stride, size_t n, size_t first_label);
    VecSimQueryReply *topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) override;
    int topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *labels, double *scores,
                  uint32_t *counts) override;
    VecSimQueryReply *rangeQuery(const void *blob, double radius, VecSimQueryParams *qp,
                                 VecSimQueryReply_Order order) override;
    VecSimBatchIterator *newBatchIterator(const void *blob, VecSimQueryParams *qp) override;
    VecSimIndexBasicInfo basicInfo() override;
    VecSimIndexDebugInfo debugInfo() override;
    VecSimIndexStatsInfo statsInfo() override;
    bool preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) override;
    void setLastSearchMode(VecSearchMode m) override { last_mode_ = m; }
    void exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) override;
    std::vector<uint8_t> preprocessQuery(const void *blob) override;
    vsgpu_store *deviceStore() override;
    void lastStats(vsgpu_stats *out) override;

    // batches formed out of concurrent single-query calls so far, and the queries they carried (see topKQuery)
    void combinerStats(size_t *batches, size_t *queries) {
        std::lock_guard<std::mutex> g(q_mu_);
        *batches = combined_batches_;
        *queries = combined_queries_;
    }

    // resolves the reference's admission/tie rule for one query from the
    // (score, id)-ordered candidates the device returned
    static void resolve(const size_t *labels, const double *scores, const uint32_t *ids, size_t cnt, size_t k,
                        std::vector<VecSimQueryResult> &out);
    // used by the batch iterator
    int allScores(const void *processed_query, std::vector<std::pair<double, size_t>> &out);

  private:
    void preprocess(const void *blob, uint8_t *out) const;
    int flush();
    // label <-> internal id. While every label equals its internal id (the common bulk-ingest case)
    // no per-row host state is kept at all; the maps are materialised on the first exception.
    bool findId(size_t label, idType *id) const;
    size_t labelOf(size_t id) const { return identity_ ? id : id_to_label_[id]; }
    void materialize();

    VecSimType type_;
    VecSimMetric metric_;
    size_t dim_, block_size_, data_size_, stored_size_;
    void *log_ctx_;
    vsgpu_store *store_ = nullptr;
    size_t count_ = 0;
    bool identity_ = true;
    hmap<size_t, idType> label_to_id_;
    hvec<size_t> id_to_label_;
    bool labels_monotone_ = true;
    size_t max_label_ = 0;
    hvec<uint8_t> pending_rows_;
    hvec<uint64_t> pending_labels_;
    VecSearchMode last_mode_ = EMPTY_MODE;
    std::mutex mu_;
    // Concurrent single-query callers (RediSearch's worker threads, vec_sim.h:147-148) are combined: the first caller to
    // arrive runs the device call for everybody who queued up meanwhile, as one batch (SURVEY §8b "Threading")
    struct PendingQuery {
        const void *blob;
        size_t k;
        VecSimQueryReply *rep;
        bool done = false;
    };
    std::mutex q_mu_;
    std::condition_variable q_cv_;
    std::vector<PendingQuery *> q_wait_;
    bool q_leader_ = false;
    size_t combined_batches_ = 0, combined_queries_ = 0;
};


// Flat, several vectors per label: vecsim_flat_multi.cpp
class FlatMultiIndex final : public VecSimIndexInterface {
  public:
    FlatMultiIndex(const BFParams &p, void *logCtx);
    ~FlatMultiIndex() override;
    bool ok() const { return store_ != nullptr; }
    int addVector(const void *blob, size_t label) override;
    long addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) override;
    int deleteVector(size_t label) override;
    double getDistanceFrom(size_t label, const void *blob) override;
    size_t indexSize() override { return id_to_label_.size(); }
    size_t indexLabelCount() override { return label_to_ids_.size(); }
    VecSimQueryReply *topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) override;
    int topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *labels, double *scores,
                  uint32_t *counts) override;
    VecSimQueryReply *rangeQuery(const void *blob, double radius, VecSimQueryParams *qp,
                                 VecSimQueryReply_Order order) override;
    VecSimBatchIterator *newBatchIterator(const void *blob, VecSimQueryParams *qp) override;
    VecSimIndexBasicInfo basicInfo() override;
    VecSimIndexDebugInfo debugInfo() override;
    VecSimIndexStatsInfo statsInfo() override;
    bool preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) override;
    void setLastSearchMode(VecSearchMode m) override { last_mode_ = m; }
    void exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) override;
    std::vector<uint8_t> preprocessQuery(const void *blob) override;
    vsgpu_store *deviceStore() override;
    void lastStats(vsgpu_stats *out) override;
    int allScores(const void *processed_query, std::vector<std::pair<double, size_t>> &out);
    static bool reduce(const uint64_t *labels, const double *scores, size_t cnt, size_t want, bool exhausted,
                       std::vector<VecSimQueryResult> &out);
    // the oldest m vectors of `label` (the tiered index moves vectors to its backend in insertion order)
    int deleteFirst(size_t label, size_t m);

  private:
    void preprocess(const void *blob, uint8_t *out) const;
    int flush();
    int removeRow(idType id);
    VecSimType type_;
    VecSimMetric metric_;
    size_t dim_, block_size_, data_size_, stored_size_;
    void *log_ctx_;
    vsgpu_store *store_ = nullptr;
    hvec<size_t> id_to_label_;
    hmap<size_t, std::vector<idType>> label_to_ids_;
    hvec<uint8_t> pending_rows_;
    hvec<uint64_t> pending_labels_;
    VecSearchMode last_mode_ = EMPTY_MODE;
    std::mutex mu_;
};

// Flat, rows sharded over several devices of one process: vecsim_flat_sharded.cpp
class ShardWorkers;
class ShardedFlatIndex final : public VecSimIndexInterface {
  public:
    ShardedFlatIndex(const BFParams &p, void *logCtx, const std::vector<int> &devices);
    ~ShardedFlatIndex() override;
    bool ok() const { return group_ != nullptr; }
    int addVector(const void *blob, size_t label) override;
    long addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) override;
    int deleteVector(size_t label) override;
    double getDistanceFrom(size_t label, const void *blob) override;
    size_t indexSize() override;
    size_t indexLabelCount() override { return indexSize(); }
    VecSimQueryReply *topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) override;
    int topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *labels, double *scores,
                  uint32_t *counts) override;
    VecSimQueryReply *rangeQuery(const void *blob, double radius, VecSimQueryParams *qp,
                                 VecSimQueryReply_Order order) override;
    VecSimBatchIterator *newBatchIterator(const void *blob, VecSimQueryParams *qp) override;
    VecSimIndexBasicInfo basicInfo() override;
    VecSimIndexDebugInfo debugInfo() override;
    VecSimIndexStatsInfo statsInfo() override;
    bool preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) override;
    void setLastSearchMode(VecSearchMode m) override { last_mode_ = m; }
    void exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) override;
    std::vector<uint8_t> preprocessQuery(const void *blob) override;
    vsgpu_store *deviceStore() override;
    void lastStats(vsgpu_stats *out) override;

    size_t shardCount() const { return shards_.size(); }
    FlatIndex *shard(size_t i) { return shards_[i].get(); }
    // rows already in the memory of one of the shards' devices: appended to that shard (labels first_label + i)
    long appendDeviceRows(const void *dev_rows, size_t stride, size_t n, size_t first_label);
    int allScores(const void *processed_query, std::vector<std::pair<double, size_t>> &out);

  private:
    size_t route(size_t label) const;
    struct Range {
        size_t first, last; // labels [first, last]
        size_t shard;
    };
    std::vector<std::unique_ptr<FlatIndex>> shards_;
    std::vector<int> devices_;
    std::vector<Range> ranges_; // bulk-ingested label ranges, sorted by first
    std::unique_ptr<ShardWorkers> workers_;
    vsgpu_group *group_ = nullptr;
    BFParams params_;
    size_t data_size_, stored_size_;
    float last_ms_ = 0.f;
    VecSearchMode last_mode_ = EMPTY_MODE;
    std::mutex mu_;
};

// HNSW: vecsim_hnsw.cpp. Single value per label (HNSWIndex_Single) or, with HNSWParams.multi, several vectors per label
// (HNSWIndex_Multi, hnsw_multi.h): queries then return each label once, with the best score among its vectors.
class HnswIndex final : public VecSimIndexInterface {
  public:
    HnswIndex(const HNSWParams &p, void *logCtx);
    ~HnswIndex() override;
    bool ok() const { return store_ != nullptr && graph_ != nullptr; }

    int addVector(const void *blob, size_t label) override;
    long addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) override;
    int deleteVector(size_t label) override;
    double getDistanceFrom(size_t label, const void *blob) override;
    size_t indexSize() override { return id_to_label_.size() - num_deleted_; }
    size_t indexLabelCount() override { return multi_ ? label_to_ids_.size() : label_to_id_.size(); }
    VecSimQueryReply *topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) override;
    int topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *labels, double *scores,
                  uint32_t *counts) override;
    VecSimQueryReply *rangeQuery(const void *blob, double radius, VecSimQueryParams *qp,
                                 VecSimQueryReply_Order order) override;
    VecSimBatchIterator *newBatchIterator(const void *blob, VecSimQueryParams *qp) override;
    VecSimIndexBasicInfo basicInfo() override;
    VecSimIndexDebugInfo debugInfo() override;
    VecSimIndexStatsInfo statsInfo() override;
    bool preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) override;
    void setLastSearchMode(VecSearchMode m) override { last_mode_ = m; }
    void exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) override;
    std::vector<uint8_t> preprocessQuery(const void *blob) override;
    vsgpu_store *deviceStore() override;
    void lastStats(vsgpu_stats *out) override;

    vsgpu_hnsw *deviceGraph();
    bool hasLabel(size_t label) {
        std::lock_guard<std::mutex> g(mu_);
        return multi_ ? label_to_ids_.count(label) != 0 : label_to_id_.count(label) != 0;
    }
    bool isMulti() const { return multi_; }
    // pushes staged vectors to the device store and graph now (the tiered index ingests from its worker threads)
    int sync() {
        std::lock_guard<std::mutex> g(mu_);
        return flush();
    }
    size_t abortPending();
    int iterNext(vsgpu_hnsw_iter *it, size_t n, size_t *labels, double *scores, size_t *count, int *depleted);
    // ids of the live rows of `label` (empty when unknown); caller holds mu_
    std::vector<idType> idsOfLocked(size_t label) const;
    void iterReset(vsgpu_hnsw_iter *it);
    void iterDestroy(vsgpu_hnsw_iter *it);
    size_t efRuntime() const { return ef_; }
    size_t M() const { return M_; }
    int elementNeighbors(size_t label, int ***out) override;
    int importGraph(const void *blobs, int processed, size_t n, const size_t *labels, const uint32_t *levels,
                    const uint32_t *l0, const uint32_t *upper, size_t upper_records, long entry, long max_level);
    // vecsim_hnsw_file.cpp: the reference's serialized index format
    int markDeletedById(idType id);
    int saveFile(const char *path);

  private:
    void preprocess(const void *blob, uint8_t *out) const;
    int flush();
    uint32_t drawLevel();
    int markDeletedLocked(idType id);

    VecSimType type_;
    VecSimMetric metric_;
    size_t dim_, block_size_, data_size_, stored_size_;
    void *log_ctx_;
    size_t M_ = 0, efc_ = 0, ef_ = 0;
    double epsilon_ = 0.01, mult_ = 0;
    std::default_random_engine level_gen_;
    vsgpu_store *store_ = nullptr;
    vsgpu_hnsw *graph_ = nullptr;
    bool multi_ = false;
    hmap<size_t, idType> label_to_id_;               // single value per label
    hmap<size_t, std::vector<idType>> label_to_ids_; // multi: the live rows of each label
    hvec<size_t> id_to_label_;
    hvec<uint8_t> id_deleted_; // tombstones, by id
    size_t num_deleted_ = 0;
    hvec<uint8_t> pending_rows_;
    hvec<uint64_t> pending_labels_;
    hvec<uint32_t> pending_levels_;
    VecSearchMode last_mode_ = EMPTY_MODE;
    std::mutex mu_;
};

// Tiered index (flat buffer in front of an HNSW backend): vecsim_tiered.cpp
class TieredBatchIterator;
class TieredIndex final : public VecSimIndexInterface {
  public:
    TieredIndex(const TieredIndexParams &p, void *logCtx);
    ~TieredIndex() override;
    bool ok() const { return front_ && back_; }
    int addVector(const void *blob, size_t label) override;
    long addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) override;
    int deleteVector(size_t label) override;
    double getDistanceFrom(size_t label, const void *blob) override;
    size_t indexSize() override;
    size_t indexLabelCount() override;
    VecSimQueryReply *topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) override;
    int topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *labels, double *scores,
                  uint32_t *counts) override;
    VecSimQueryReply *rangeQuery(const void *blob, double radius, VecSimQueryParams *qp,
                                 VecSimQueryReply_Order order) override;
    VecSimBatchIterator *newBatchIterator(const void *blob, VecSimQueryParams *qp) override;
    VecSimIndexBasicInfo basicInfo() override;
    VecSimIndexDebugInfo debugInfo() override;
    VecSimIndexStatsInfo statsInfo() override;
    bool preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) override;
    void setLastSearchMode(VecSearchMode m) override;
    void exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) override;
    std::vector<uint8_t> preprocessQuery(const void *blob) override;
    vsgpu_store *deviceStore() override;
    void lastStats(vsgpu_stats *out) override;
    int elementNeighbors(size_t label, int ***out) override;

    VecSimIndexInterface *frontend() { return front_.get(); }
    HnswIndex *backend() { return back_.get(); }
    void acquireSharedLocks();
    void releaseSharedLocks();
    size_t pendingJobs() {
        std::shared_lock<std::shared_mutex> flat(flat_guard_);
        return pending_.size();
    }

  private:
    friend class TieredBatchIterator;
    static void executeJobWrapper(AsyncJob *job);
    void executeInsertJob(AsyncJob *job);
    void invalidateJobLocked(size_t label);

    std::unique_ptr<VecSimIndexInterface> front_; // FlatIndex, or FlatMultiIndex over a multi-value backend
    std::unique_ptr<HnswIndex> back_;
    bool multi_ = false;
    void *job_queue_, *job_queue_ctx_;
    SubmitCB submit_;
    size_t flat_limit_, swap_threshold_;
    size_t data_size_ = 0;
    std::shared_mutex flat_guard_, main_guard_; // always taken in this order
    std::mutex drain_mu_;
    std::unordered_map<size_t, std::vector<AsyncJob *>> label_to_job_; // the pending job(s) of each label in the flat buffer
    std::vector<AsyncJob *> pending_;                      // the same jobs in submission order
    std::vector<AsyncJob *> parked_;                       // vectors the backend refused three times: they stay in the buffer
    std::atomic<size_t> direct_insertions_{0};
    std::shared_ptr<std::atomic<bool>> alive_;
};

VecSimBatchIterator *new_flat_batch_iterator(std::function<int(const void *, std::vector<std::pair<double, size_t>> &)> score_all,
                                             size_t label_count, std::vector<uint8_t> query, void *tctx);
VecSimIndexInterface *load_hnsw_file(const char *path, std::string &err);
size_t tiered_merge_for_test(const size_t *a_ids, const double *a_scores, size_t na, const size_t *b_ids, const double *b_scores,
                             size_t nb, size_t limit, size_t *out_ids, double *out_scores, size_t *taken);

} // namespace vsb
