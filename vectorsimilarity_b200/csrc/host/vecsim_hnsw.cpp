// HNSW index: host bookkeeping of the reference's HNSWIndex_Single
// (/root/reference/src/VecSim/algorithms/hnsw/hnsw.h:1616-1650 ctor, :418-422 level draw, :1860-1960
// store + index a vector, :2037-2084 topKQuery, :2152-2186 rangeQuery; hnsw_single.h:134-170 add /
// delete) over a device-resident row store and graph. Every distance, the traversals and the graph
// construction run in libvsgpu.so (vsgpu_hnsw_*); there is no CPU fallback.
#include "vecsim_index.h"
#include "vecsim_numeric.h"
#include "vecsim_hybrid.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>
#include <unordered_set>

namespace vsb {

static constexpr size_t HNSW_FLUSH_ROWS = 4096;

HnswIndex::HnswIndex(const HNSWParams &p, void *logCtx)
    : type_(p.type), metric_(p.metric), dim_(p.dim), block_size_(p.blockSize ? p.blockSize : 1024),
      data_size_(type_size(p.type) * p.dim), stored_size_(stored_size(p.type, p.dim, p.metric)), log_ctx_(logCtx) {
    M_ = p.M ? p.M : 16;                                              // HNSW_DEFAULT_M
    efc_ = std::max<size_t>(p.efConstruction ? p.efConstruction : 200, M_); // hnsw.h:1632-1633
    ef_ = p.efRuntime ? p.efRuntime : 10;
    epsilon_ = p.epsilon > 0.0 ? p.epsilon : 0.01;
    if (M_ <= 1 || 2 * M_ > 512) return; // the reference throws for M<=1 (hnsw.h:1644); >256 is a device limit
    mult_ = 1.0 / std::log(1.0 * (double)M_);
    level_gen_.seed(100); // hnsw.h:230 default random_seed
    multi_ = p.multi;
    store_ = vsgpu_store_create(globals().device, (int)type_, (int)metric_, dim_, p.initialCapacity);
    if (store_) graph_ = vsgpu_hnsw_create(store_, M_, efc_);
    if (graph_) vsgpu_hnsw_set_multi(graph_, multi_ ? 1 : 0);
}

std::vector<idType> HnswIndex::idsOfLocked(size_t label) const {
    if (multi_) {
        auto it = label_to_ids_.find(label);
        return it == label_to_ids_.end() ? std::vector<idType>() : it->second;
    }
    auto it = label_to_id_.find(label);
    return it == label_to_id_.end() ? std::vector<idType>() : std::vector<idType>{it->second};
}

HnswIndex::~HnswIndex() {
    if (graph_) vsgpu_hnsw_destroy(graph_);
    if (store_) vsgpu_store_destroy(store_);
}

void HnswIndex::preprocess(const void *blob, uint8_t *out) const {
    std::memcpy(out, blob, data_size_);
    if (metric_ == VecSimMetric_Cosine) normalize_blob(out, dim_, type_);
}

std::vector<uint8_t> HnswIndex::preprocessQuery(const void *blob) {
    std::vector<uint8_t> q(stored_size_);
    preprocess(blob, q.data());
    return q;
}

// getRandomLevel (hnsw.h:418-422): the same engine and distribution objects as the reference, so
// the level sequence is the reference's for the same insertion order.
uint32_t HnswIndex::drawLevel() {
    std::uniform_real_distribution<double> distribution(0.0, 1.0);
    const double r = -std::log(distribution(level_gen_)) * mult_;
    return (uint32_t)(size_t)r;
}

int HnswIndex::flush() {
    if (pending_labels_.empty()) return 0;
    const size_t n = pending_labels_.size();
    int rc = vsgpu_store_append(store_, pending_rows_.data(), stored_size_, pending_labels_.data(), n);
    if (rc != VSGPU_OK) return rc; // nothing appended: the vectors stay staged
    rc = vsgpu_hnsw_insert(graph_, n, pending_levels_.data());
    if (rc != VSGPU_OK) {
        // the graph did not take them: roll the store back so row ids and node ids stay in step; the vectors stay staged
        // (the caller may retry, or drop them with abortPending)
        vsgpu_store_truncate(store_, vsgpu_hnsw_size(graph_));
        return rc;
    }
    pending_rows_.clear();
    pending_labels_.clear();
    pending_levels_.clear();
    return rc;
}

// Forget the vectors that are staged but not on the device (after a failed flush): their ids are the tail of the id space.
// Labels whose staged vector replaced an older one stay deleted — the caller still holds the new vector.
size_t HnswIndex::abortPending() {
    std::lock_guard<std::mutex> g(mu_);
    const size_t n = pending_labels_.size();
    for (size_t i = 0; i < n; i++) {
        const size_t id = id_to_label_.size() - 1;
        if (multi_) {
            auto it = label_to_ids_.find(id_to_label_[id]);
            if (it != label_to_ids_.end()) {
                auto &v = it->second;
                v.erase(std::remove(v.begin(), v.end(), (idType)id), v.end());
                if (v.empty()) label_to_ids_.erase(it);
            }
        } else {
            auto it = label_to_id_.find(id_to_label_[id]);
            if (it != label_to_id_.end() && it->second == (idType)id) label_to_id_.erase(it);
        }
        id_to_label_.pop_back();
        id_deleted_.pop_back();
    }
    pending_rows_.clear();
    pending_labels_.clear();
    pending_levels_.clear();
    return n;
}

int HnswIndex::addVector(const void *blob, size_t label) {
    std::lock_guard<std::mutex> g(mu_);
    int ret = 1;
    if (!multi_) {
        auto it = label_to_id_.find(label);
        if (it != label_to_id_.end()) {
            // hnsw_single.h:153-163 deletes the old vector and appends the new one. The old node is
            // tombstoned here instead of being cut out of the graph.
            if (markDeletedLocked(it->second) != 0) return -1;
            label_to_id_.erase(it);
            ret = 0;
        }
    }
    const size_t id = id_to_label_.size();
    if (id >= 0xfffffffeull) return -1;
    id_to_label_.push_back(label);
    id_deleted_.push_back(0);
    if (multi_) label_to_ids_[label].push_back((idType)id); // hnsw_multi.h:186-196: another vector under the label, always new
    else label_to_id_[label] = (idType)id;
    const size_t off = pending_rows_.size();
    pending_rows_.resize(off + stored_size_);
    preprocess(blob, pending_rows_.data() + off);
    pending_labels_.push_back(label);
    pending_levels_.push_back(drawLevel());
    if (pending_labels_.size() >= HNSW_FLUSH_ROWS)
        if (flush() != 0) return -1;
    return ret;
}

long HnswIndex::addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) {
    long added = 0;
    for (size_t i = 0; i < n; i++) {
        const int rc = addVector((const uint8_t *)blobs + i * data_size_, labels ? labels[i] : first_label + i);
        if (rc < 0) return -1;
        added += rc;
    }
    return added;
}

int HnswIndex::markDeletedLocked(idType id) {
    if (flush() != 0) return -1;
    if (id < id_deleted_.size() && id_deleted_[id]) return 0;
    if (vsgpu_hnsw_set_deleted(graph_, id, 1) != VSGPU_OK) return -1;
    if (id < id_deleted_.size()) id_deleted_[id] = 1;
    num_deleted_++;
    return 0;
}

int HnswIndex::deleteVector(size_t label) {
    std::lock_guard<std::mutex> g(mu_);
    if (multi_) { // hnsw_multi.h:221-247: every vector of the label
        auto it = label_to_ids_.find(label);
        if (it == label_to_ids_.end()) return 0;
        int n = 0;
        for (idType id : it->second)
            if (markDeletedLocked(id) == 0) n++;
        label_to_ids_.erase(it);
        return n;
    }
    auto it = label_to_id_.find(label);
    if (it == label_to_id_.end()) return 0;
    if (markDeletedLocked(it->second) != 0) return 0;
    label_to_id_.erase(it);
    return 1;
}

double HnswIndex::getDistanceFrom(size_t label, const void *blob) {
    std::lock_guard<std::mutex> g(mu_);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    const std::vector<idType> ids = idsOfLocked(label);
    if (ids.empty()) return nan;
    if (flush() != 0) return nan;
    std::vector<double> d(ids.size());
    if (vsgpu_distances(store_, blob, ids.data(), ids.size(), d.data()) != VSGPU_OK) return nan;
    double best = d[0]; // hnsw_multi.h:81-95: the closest of the label's vectors
    for (double v : d) best = (best < v) ? best : v;
    return best;
}

void HnswIndex::exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) {
    std::lock_guard<std::mutex> g(mu_);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (size_t i = 0; i < n; i++) out[i] = nan;
    if (flush() != 0) return;
    std::vector<uint32_t> ids;
    std::vector<size_t> pos;
    for (size_t i = 0; i < n; i++)
        for (idType id : idsOfLocked(labels[i])) {
            ids.push_back(id);
            pos.push_back(i);
        }
    if (ids.empty()) return;
    std::vector<double> d(ids.size());
    if (vsgpu_distances(store_, processed_query, ids.data(), ids.size(), d.data()) != VSGPU_OK) return;
    for (size_t j = 0; j < ids.size(); j++)
        if (!(out[pos[j]] <= d[j])) out[pos[j]] = d[j]; // minimum over a label's vectors (NaN = not set yet)
}

int HnswIndex::topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *out_labels,
                         double *out_scores, uint32_t *out_counts) {
    std::lock_guard<std::mutex> g(mu_);
    void *tctx = qp ? qp->timeoutCtx : nullptr;
    last_mode_ = STANDARD_KNN;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    auto pad_all = [&]() {
        for (size_t i = 0; i < nq * k; i++) {
            if (out_labels) out_labels[i] = (size_t)-1;
            if (out_scores) out_scores[i] = nan;
        }
        if (out_counts)
            for (size_t q = 0; q < nq; q++) out_counts[q] = 0;
    };
    if (nq == 0) return 0;
    if (k == 0 || id_to_label_.empty()) {
        pad_all();
        return 0;
    }
    if (timed_out(tctx)) {
        pad_all();
        return 1;
    }
    if (flush() != 0) return -1;
    size_t ef = ef_;
    if (qp && qp->hnswRuntimeParams.efRuntime != 0) ef = qp->hnswRuntimeParams.efRuntime;
    std::vector<uint8_t> qbuf(nq * stored_size_);
    for (size_t q = 0; q < nq; q++) preprocess((const uint8_t *)queries + q * data_size_, qbuf.data() + q * stored_size_);
    static_assert(sizeof(size_t) == sizeof(uint64_t), "labelType is 64-bit");
    std::vector<uint32_t> counts(nq);
    const int rc = vsgpu_hnsw_topk(graph_, qbuf.data(), nq, stored_size_, k, ef, (uint64_t *)out_labels, out_scores, nullptr,
                                   counts.data());
    if (rc != VSGPU_OK) return -1;
    if (timed_out(tctx)) {
        pad_all();
        return 1;
    }
    if (out_counts) std::copy(counts.begin(), counts.end(), out_counts);
    return 0;
}

VecSimQueryReply *HnswIndex::topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) {
    auto *rep = new VecSimQueryReply();
    if (k == 0 || indexSize() == 0) {
        last_mode_ = STANDARD_KNN;
        return rep;
    }
    const size_t cap = std::min(k, multi_ ? indexLabelCount() : id_to_label_.size());
    std::vector<size_t> labels(cap);
    std::vector<double> scores(cap);
    uint32_t cnt = 0;
    const int rc = topKBatch(blob, 1, cap, qp, labels.data(), scores.data(), &cnt);
    if (rc == 1) rep->code = VecSim_QueryReply_TimedOut;
    if (rc != 0) return rep;
    rep->results.resize(cnt);
    for (uint32_t i = 0; i < cnt; i++) rep->results[i] = {labels[i], scores[i]};
    return rep;
}

VecSimQueryReply *HnswIndex::rangeQuery(const void *blob, double radius, VecSimQueryParams *qp, VecSimQueryReply_Order order) {
    std::lock_guard<std::mutex> g(mu_);
    auto *rep = new VecSimQueryReply();
    void *tctx = qp ? qp->timeoutCtx : nullptr;
    last_mode_ = RANGE_QUERY;
    if (id_to_label_.empty()) return rep;
    if (timed_out(tctx)) {
        rep->code = VecSim_QueryReply_TimedOut;
        return rep;
    }
    if (flush() != 0) return rep;
    double eps = epsilon_;
    if (qp && qp->hnswRuntimeParams.epsilon != 0.0) eps = qp->hnswRuntimeParams.epsilon;
    std::vector<uint8_t> q(stored_size_);
    preprocess(blob, q.data());
    size_t cap = 4096, count = 0;
    std::vector<uint64_t> lab;
    std::vector<double> sc;
    for (int attempt = 0; attempt < 2; attempt++) {
        lab.resize(cap);
        sc.resize(cap);
        const int rc = vsgpu_hnsw_range(graph_, q.data(), radius, eps, cap, lab.data(), sc.data(), nullptr, &count);
        if (rc == VSGPU_OK) break;
        if (rc != VSGPU_ERR_OVERFLOW) return rep;
        cap = count;
    }
    if (timed_out(tctx)) {
        rep->code = VecSim_QueryReply_TimedOut;
        return rep;
    }
    rep->results.resize(count);
    for (size_t i = 0; i < count; i++) rep->results[i] = {(size_t)lab[i], sc[i]};
    if (multi_ && count > 1) {
        // unique_results_container (utils/query_result_utils.h, hnsw_multi.h:73-79): each label once, with its best score
        std::sort(rep->results.begin(), rep->results.end(), [](const VecSimQueryResult &a, const VecSimQueryResult &b) {
            if (a.id != b.id) return a.id < b.id;
            return a.score < b.score;
        });
        rep->results.erase(std::unique(rep->results.begin(), rep->results.end(),
                                       [](const VecSimQueryResult &a, const VecSimQueryResult &b) { return a.id == b.id; }),
                           rep->results.end());
    }
    if (order == BY_ID)
        std::sort(rep->results.begin(), rep->results.end(),
                  [](const VecSimQueryResult &a, const VecSimQueryResult &b) { return a.id < b.id; });
    else
        std::sort(rep->results.begin(), rep->results.end(), [](const VecSimQueryResult &a, const VecSimQueryResult &b) {
            if (a.score < b.score) return true;
            if (b.score < a.score) return false;
            return a.id < b.id;
        });
    return rep;
}

// HNSW_BatchIterator (hnsw_batch_iterator.h:59-267): the traversal state stays on the device between
// calls (vsgpu_hnsw_iter_*), each Next is one resumed scan with the reference's heap rules.
class HnswBatchIterator final : public VecSimBatchIterator {
  public:
    HnswBatchIterator(HnswIndex *idx, vsgpu_hnsw_iter *it, void *tctx) : idx_(idx), it_(it), tctx_(tctx) {}
    ~HnswBatchIterator() override { idx_->iterDestroy(it_); }
    VecSimQueryReply *next(size_t n, VecSimQueryReply_Order order) override {
        auto *rep = new VecSimQueryReply();
        if (!it_) {
            depleted_ = true;
            return rep;
        }
        if (timed_out(tctx_)) {
            rep->code = VecSim_QueryReply_TimedOut;
            return rep;
        }
        std::vector<size_t> labels(std::max<size_t>(n, 1));
        std::vector<double> scores(std::max<size_t>(n, 1));
        size_t count = 0;
        int dep = 0;
        if (idx_->iterNext(it_, n, labels.data(), scores.data(), &count, &dep) != 0) return rep;
        depleted_ = dep != 0;
        rep->results.resize(count);
        for (size_t i = 0; i < count; i++) rep->results[i] = {labels[i], scores[i]};
        if (order == BY_ID)
            std::sort(rep->results.begin(), rep->results.end(),
                      [](const VecSimQueryResult &a, const VecSimQueryResult &b) { return a.id < b.id; });
        return rep;
    }
    bool hasNext() override { return !depleted_; }
    void reset() override {
        if (it_) idx_->iterReset(it_);
        depleted_ = false;
    }

  private:
    HnswIndex *idx_;
    vsgpu_hnsw_iter *it_;
    void *tctx_;
    bool depleted_ = false;
};

VecSimBatchIterator *HnswIndex::newBatchIterator(const void *blob, VecSimQueryParams *qp) {
    std::lock_guard<std::mutex> g(mu_);
    vsgpu_hnsw_iter *it = nullptr;
    if (flush() == 0) {
        std::vector<uint8_t> q(stored_size_);
        preprocess(blob, q.data());
        size_t ef = ef_;
        if (qp && qp->hnswRuntimeParams.efRuntime > 0) ef = qp->hnswRuntimeParams.efRuntime;
        it = vsgpu_hnsw_iter_create(graph_, q.data(), ef);
    }
    return new HnswBatchIterator(this, it, qp ? qp->timeoutCtx : nullptr);
}

int HnswIndex::iterNext(vsgpu_hnsw_iter *it, size_t n, size_t *labels, double *scores, size_t *count, int *depleted) {
    std::lock_guard<std::mutex> g(mu_);
    if (flush() != 0) return -1;
    static_assert(sizeof(size_t) == sizeof(uint64_t), "labelType is 64-bit");
    if (vsgpu_hnsw_iter_next(it, n, multi_ ? label_to_ids_.size() : label_to_id_.size(), (uint64_t *)labels, scores, nullptr, count,
                             depleted) != VSGPU_OK)
        return -1;
    if (multi_) { // HNSWMulti_BatchIterator::returned: later batches skip every vector of the labels handed out now
        std::vector<uint32_t> ids;
        for (size_t i = 0; i < *count; i++)
            for (idType id : idsOfLocked(labels[i])) ids.push_back(id);
        if (vsgpu_hnsw_iter_mark_returned(it, ids.data(), ids.size()) != VSGPU_OK) return -1;
    }
    return 0;
}
void HnswIndex::iterReset(vsgpu_hnsw_iter *it) {
    std::lock_guard<std::mutex> g(mu_);
    vsgpu_hnsw_iter_reset(it);
}
void HnswIndex::iterDestroy(vsgpu_hnsw_iter *it) {
    std::lock_guard<std::mutex> g(mu_);
    vsgpu_hnsw_iter_destroy(it);
}

VecSimIndexBasicInfo HnswIndex::basicInfo() {
    VecSimIndexBasicInfo b{};
    b.algo = VecSimAlgo_HNSWLIB;
    b.metric = metric_;
    b.type = type_;
    b.isMulti = multi_;
    b.isTiered = false;
    b.isDisk = false;
    b.blockSize = block_size_;
    b.dim = dim_;
    return b;
}

VecSimIndexStatsInfo HnswIndex::statsInfo() {
    VecSimIndexStatsInfo s{};
    s.memory = sizeof(*this) + id_to_label_.capacity() * (sizeof(size_t) + 1) + (label_to_id_.size() + label_to_ids_.size()) * 32 +
               (multi_ ? id_to_label_.size() * sizeof(idType) : 0) + pending_rows_.capacity() +
               (store_ ? vsgpu_store_device_bytes(store_) : 0) + (graph_ ? vsgpu_hnsw_device_bytes(graph_) : 0);
    s.numberOfMarkedDeleted = num_deleted_;
    return s;
}

VecSimIndexDebugInfo HnswIndex::debugInfo() {
    std::lock_guard<std::mutex> g(mu_);
    flush();
    VecSimIndexDebugInfo d{};
    d.commonInfo.basicInfo = basicInfo();
    d.commonInfo.indexSize = id_to_label_.size() - num_deleted_;
    d.commonInfo.indexLabelCount = multi_ ? label_to_ids_.size() : label_to_id_.size();
    d.commonInfo.memory = statsInfo().memory;
    d.commonInfo.lastMode = last_mode_;
    long ep = -1, ml = -1;
    vsgpu_hnsw_entry(graph_, &ep, &ml);
    d.hnswInfo.M = M_;
    d.hnswInfo.efConstruction = efc_;
    d.hnswInfo.efRuntime = ef_;
    d.hnswInfo.epsilon = epsilon_;
    d.hnswInfo.max_level = (size_t)ml; // HNSW_INVALID_LEVEL == SIZE_MAX when empty
    d.hnswInfo.entrypoint = ep < 0 ? (size_t)-1 : id_to_label_[(size_t)ep];
    d.hnswInfo.visitedNodesPoolSize = 1;
    d.hnswInfo.numberOfMarkedDeletedNodes = num_deleted_;
    return d;
}

// hnsw.h:2340-2400 uses a CPU-fitted tree; here a cost comparison with device rates (vecsim_hybrid.h):
// a batch pass must surface ~k / r results at one latency-bound hop each, an ad-hoc pass gathers the subset's rows.
bool HnswIndex::preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) {
    const bool res = prefer_adhoc_hnsw(indexSize(), stored_size_, subsetSize, k);
    last_mode_ = res ? (initial_check ? HYBRID_ADHOC_BF : HYBRID_BATCHES_TO_ADHOC_BF) : HYBRID_BATCHES;
    return res;
}

vsgpu_store *HnswIndex::deviceStore() {
    std::lock_guard<std::mutex> g(mu_);
    if (flush() != 0) return nullptr;
    return store_;
}

vsgpu_hnsw *HnswIndex::deviceGraph() {
    std::lock_guard<std::mutex> g(mu_);
    if (flush() != 0) return nullptr;
    return graph_;
}

void HnswIndex::lastStats(vsgpu_stats *out) {
    *out = vsgpu_stats{};
    unsigned long long ev = 0, hops = 0;
    float ms = 0;
    vsgpu_hnsw_last_stats(graph_, &ev, &hops, &ms);
    out->path = 2;
    out->kernel_launches = 1;
    out->candidates = ev;
    out->scan_ms = ms;
    out->total_ms = ms;
}

// getHNSWElementNeighbors (hnsw.h:2414-2441): per level [count, neighbour labels...], NULL-terminated
int HnswIndex::elementNeighbors(size_t label, int ***out) {
    std::lock_guard<std::mutex> g(mu_);
    *out = nullptr;
    if (multi_) return VecSimDebugCommandCode_MultiNotSupported; // hnsw_multi.h: one label has several nodes
    auto it = label_to_id_.find(label);
    if (it == label_to_id_.end()) return VecSimDebugCommandCode_LabelNotExists;
    if (flush() != 0) return VecSimDebugCommandCode_BadIndex;
    uint32_t lvl = 0;
    if (vsgpu_hnsw_node(graph_, it->second, &lvl, nullptr, 0) != VSGPU_OK) return VecSimDebugCommandCode_BadIndex;
    std::vector<uint32_t> rec((2 * M_ + 1) + (size_t)lvl * (M_ + 1));
    if (vsgpu_hnsw_node(graph_, it->second, &lvl, rec.data(), rec.size()) != VSGPU_OK) return VecSimDebugCommandCode_BadIndex;
    int **res = new int *[lvl + 2];
    for (uint32_t l = 0; l <= lvl; l++) {
        const uint32_t *r = l == 0 ? rec.data() : rec.data() + (2 * M_ + 1) + (size_t)(l - 1) * (M_ + 1);
        res[l] = new int[r[0] + 1];
        res[l][0] = (int)r[0];
        for (uint32_t i = 0; i < r[0]; i++) res[l][i + 1] = (int)id_to_label_[r[1 + i]];
    }
    res[lvl + 1] = nullptr;
    *out = res;
    return VecSimDebugCommandCode_OK;
}

// Adopt a graph built elsewhere over rows in insertion order (bulk load; the serialized-file reader of
// SURVEY §8 row f3 lands on this).
int HnswIndex::importGraph(const void *blobs, int processed, size_t n, const size_t *labels, const uint32_t *levels,
                           const uint32_t *l0, const uint32_t *upper, size_t upper_records, long entry, long max_level) {
    std::lock_guard<std::mutex> g(mu_);
    if (!id_to_label_.empty()) return -1;
    std::vector<uint8_t> rows(n * stored_size_);
    std::vector<uint64_t> lab(n);
    for (size_t i = 0; i < n; i++) {
        if (processed) std::memcpy(rows.data() + i * stored_size_, (const uint8_t *)blobs + i * stored_size_, stored_size_);
        else preprocess((const uint8_t *)blobs + i * data_size_, rows.data() + i * stored_size_);
        lab[i] = labels ? labels[i] : i;
    }
    if (vsgpu_store_append(store_, rows.data(), stored_size_, lab.data(), n) != VSGPU_OK) return -1;
    if (vsgpu_hnsw_import(graph_, n, levels, l0, upper, upper_records, entry, max_level) != VSGPU_OK) return -1;
    id_to_label_.assign(lab.begin(), lab.end());
    id_deleted_.assign(n, 0);
    for (size_t i = 0; i < n; i++) {
        if (multi_) label_to_ids_[lab[i]].push_back((idType)i);
        else label_to_id_[lab[i]] = (idType)i;
    }
    return 0;
}

} // namespace vsb
