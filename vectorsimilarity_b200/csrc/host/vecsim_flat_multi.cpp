// Multi-value flat index: host bookkeeping of the reference's BruteForceIndex_Multi
// (/root/reference/src/VecSim/algorithms/brute_force/brute_force_multi.h:97-250, bfm_batch_iterator.h:24-53,
// utils/updatable_heap.h:24-111): several vectors per label, a label's score is the minimum over its vectors,
// replies hold each label once. Rows live in HBM (vsgpu_store); all distances and the row-level selection run in
// libvsgpu.so, the per-label reduction of the (few) selected rows is host code. No CPU fallback.
#include "vecsim_index.h"
#include "vecsim_hybrid.h"
#include "vecsim_numeric.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <unordered_set>

namespace vsb {

static constexpr size_t FLUSH_ROWS = 8192;

FlatMultiIndex::FlatMultiIndex(const BFParams &p, void *logCtx)
    : type_(p.type), metric_(p.metric), dim_(p.dim), block_size_(p.blockSize ? p.blockSize : 1024),
      data_size_(type_size(p.type) * p.dim), stored_size_(stored_size(p.type, p.dim, p.metric)), log_ctx_(logCtx) {
    store_ = vsgpu_store_create(globals().device, (int)type_, (int)metric_, dim_, p.initialCapacity);
}

FlatMultiIndex::~FlatMultiIndex() {
    if (store_) vsgpu_store_destroy(store_);
}

void FlatMultiIndex::preprocess(const void *blob, uint8_t *out) const {
    std::memcpy(out, blob, data_size_);
    if (metric_ == VecSimMetric_Cosine) normalize_blob(out, dim_, type_);
}

std::vector<uint8_t> FlatMultiIndex::preprocessQuery(const void *blob) {
    std::vector<uint8_t> q(stored_size_);
    preprocess(blob, q.data());
    return q;
}

int FlatMultiIndex::flush() {
    if (pending_labels_.empty()) return 0;
    const int rc = vsgpu_store_append(store_, pending_rows_.data(), stored_size_, pending_labels_.data(), pending_labels_.size());
    if (rc != VSGPU_OK) return rc;
    pending_rows_.clear();
    pending_labels_.clear();
    return 0;
}

vsgpu_store *FlatMultiIndex::deviceStore() {
    std::lock_guard<std::mutex> g(mu_);
    return flush() == 0 ? store_ : nullptr;
}
void FlatMultiIndex::lastStats(vsgpu_stats *out) { vsgpu_last_stats(store_, out); }

// brute_force_multi.h:143-147: always a new vector
int FlatMultiIndex::addVector(const void *blob, size_t label) {
    std::lock_guard<std::mutex> g(mu_);
    const size_t id = id_to_label_.size();
    if (id >= 0xfffffffeull) return -1;
    id_to_label_.push_back(label);
    label_to_ids_[label].push_back((idType)id);
    const size_t off = pending_rows_.size();
    pending_rows_.resize(off + stored_size_);
    preprocess(blob, pending_rows_.data() + off);
    pending_labels_.push_back(label);
    if (pending_labels_.size() >= FLUSH_ROWS)
        if (flush() != 0) return -1;
    return 1;
}

long FlatMultiIndex::addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) {
    long added = 0;
    for (size_t i = 0; i < n; i++) {
        const int rc = addVector((const uint8_t *)blobs + i * data_size_, labels ? labels[i] : first_label + i);
        if (rc < 0) return -1;
        added += rc;
    }
    return added;
}

// removeVector (brute_force.h:195-224): the last row moves into the hole; the moved row's label learns its new id
int FlatMultiIndex::removeRow(idType id) {
    const idType last = (idType)(id_to_label_.size() - 1);
    if (vsgpu_store_remove_swap(store_, id) != VSGPU_OK) return -1;
    if (id != last) {
        const size_t moved = id_to_label_[last];
        id_to_label_[id] = moved;
        // back to front, as replaceIdOfLabel does (brute_force_multi.h:244-263): while deleteVector walks a label's own
        // list, the entries before its cursor are stale (already removed) and may equal `last` by coincidence
        std::vector<idType> &ids = label_to_ids_[moved];
        for (size_t t = ids.size(); t-- > 0;)
            if (ids[t] == last) {
                ids[t] = id;
                break;
            }
    }
    id_to_label_.pop_back();
    return 0;
}

// brute_force_multi.h:149-163: every vector of the label, in the order of its id list (the list is patched in place
// when one of its own rows is the one that moves)
int FlatMultiIndex::deleteVector(size_t label) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = label_to_ids_.find(label);
    if (it == label_to_ids_.end()) return 0;
    if (flush() != 0) return 0;
    int ret = 0;
    std::vector<idType> &ids = it->second;
    for (size_t i = 0; i < ids.size(); i++) {
        if (removeRow(ids[i]) != 0) break;
        ret++;
    }
    label_to_ids_.erase(label);
    return ret;
}

int FlatMultiIndex::deleteFirst(size_t label, size_t m) {
    std::lock_guard<std::mutex> g(mu_);
    auto it = label_to_ids_.find(label);
    if (it == label_to_ids_.end() || m == 0) return 0;
    if (flush() != 0) return 0;
    std::vector<idType> &ids = it->second; // insertion order; patched in place when one of its own rows moves
    size_t done = 0;
    for (; done < m && done < ids.size(); done++)
        if (removeRow(ids[done]) != 0) break;
    ids.erase(ids.begin(), ids.begin() + (ptrdiff_t)done);
    if (ids.empty()) label_to_ids_.erase(it);
    return (int)done;
}

double FlatMultiIndex::getDistanceFrom(size_t label, const void *blob) {
    std::lock_guard<std::mutex> g(mu_);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    auto it = label_to_ids_.find(label);
    if (it == label_to_ids_.end()) return nan;
    if (flush() != 0) return nan;
    std::vector<uint32_t> ids(it->second.begin(), it->second.end());
    std::vector<double> d(ids.size());
    if (vsgpu_distances(store_, blob, ids.data(), ids.size(), d.data()) != VSGPU_OK) return nan;
    double best = std::numeric_limits<double>::infinity(); // brute_force_multi.h:233-237: dist = (dist < d) ? dist : d
    for (double v : d) best = (best < v) ? best : v;
    return best;
}

void FlatMultiIndex::exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) {
    std::lock_guard<std::mutex> g(mu_);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (size_t i = 0; i < n; i++) out[i] = nan;
    if (flush() != 0) return;
    std::vector<uint32_t> ids;
    std::vector<size_t> owner;
    for (size_t i = 0; i < n; i++) {
        auto it = label_to_ids_.find(labels[i]);
        if (it == label_to_ids_.end()) continue;
        for (idType id : it->second) {
            ids.push_back(id);
            owner.push_back(i);
        }
    }
    if (ids.empty()) return;
    std::vector<double> d(ids.size());
    if (vsgpu_distances(store_, processed_query, ids.data(), ids.size(), d.data()) != VSGPU_OK) return;
    for (size_t j = 0; j < ids.size(); j++) {
        double &o = out[owner[j]];
        o = (o == o && o < d[j]) ? o : d[j];
    }
}

// Per-label reduction of rows sorted ascending by (score, id): each label keeps its first (= minimum) row; the reply
// is ascending (score, label) like the drained updatable_max_heap (brute_force.h:284-288, updatable_heap.h:66-76).
// Returns false when fewer than `want` labels were found although more rows exist (caller widens the selection).
bool FlatMultiIndex::reduce(const uint64_t *labels, const double *scores, size_t cnt, size_t want, bool exhausted,
                            std::vector<VecSimQueryResult> &out) {
    out.clear();
    std::unordered_set<size_t> seen;
    seen.reserve(cnt * 2);
    size_t i = 0;
    for (; i < cnt; i++) {
        // labels appear in non-decreasing order of their minimum score; once `want` are known, keep going only through
        // the tie group of the want-th score (other labels with the same best score compete by label)
        if (out.size() >= want && scores[i] > out[want - 1].score) break;
        if (seen.insert((size_t)labels[i]).second) out.push_back({(size_t)labels[i], scores[i]});
    }
    if (i == cnt && !exhausted) return false; // ran out of selected rows before the outcome was decided: widen
    std::sort(out.begin(), out.end(), [](const VecSimQueryResult &a, const VecSimQueryResult &b) {
        if (a.score < b.score) return true;
        if (b.score < a.score) return false;
        return a.id < b.id;
    });
    if (out.size() > want) out.resize(want);
    return true;
}

int FlatMultiIndex::topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *out_labels,
                              double *out_scores, uint32_t *out_counts) {
    std::lock_guard<std::mutex> g(mu_);
    void *tctx = qp ? qp->timeoutCtx : nullptr;
    last_mode_ = STANDARD_KNN;
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (size_t i = 0; i < nq * k; i++) {
        if (out_labels) out_labels[i] = (size_t)-1;
        if (out_scores) out_scores[i] = nan;
    }
    if (out_counts)
        for (size_t q = 0; q < nq; q++) out_counts[q] = 0;
    if (nq == 0 || k == 0 || id_to_label_.empty()) return 0;
    if (timed_out(tctx)) return 1;
    if (flush() != 0) return -1;
    const size_t n = id_to_label_.size();
    const size_t want = std::min(k, label_to_ids_.size());
    std::vector<uint8_t> qbuf(nq * stored_size_);
    for (size_t q = 0; q < nq; q++) preprocess((const uint8_t *)queries + q * data_size_, qbuf.data() + q * stored_size_);
    // rows per label on average decide the first selection width; widen (x4) for the queries that come up short
    const double mult = (double)n / (double)label_to_ids_.size();
    size_t k_sel = std::min(n, (size_t)std::ceil((double)want * std::max(1.0, mult) * 1.5) + 8);
    std::vector<size_t> todo(nq);
    for (size_t q = 0; q < nq; q++) todo[q] = q;
    std::vector<VecSimQueryResult> res;
    while (!todo.empty()) {
        const size_t m = todo.size();
        std::vector<uint8_t> qsub(m * stored_size_);
        for (size_t i = 0; i < m; i++) std::memcpy(qsub.data() + i * stored_size_, qbuf.data() + todo[i] * stored_size_, stored_size_);
        std::vector<uint64_t> lab(m * k_sel);
        std::vector<double> sc(m * k_sel);
        if (vsgpu_topk(store_, qsub.data(), m, stored_size_, k_sel, (unsigned)globals().topk_mode, lab.data(), sc.data(), nullptr,
                       nullptr) != VSGPU_OK)
            return -1;
        std::vector<size_t> again;
        for (size_t i = 0; i < m; i++) {
            const size_t q = todo[i];
            if (!reduce(lab.data() + i * k_sel, sc.data() + i * k_sel, k_sel, want, k_sel >= n, res)) {
                again.push_back(q);
                continue;
            }
            for (size_t j = 0; j < res.size(); j++) {
                if (out_labels) out_labels[q * k + j] = res[j].id;
                if (out_scores) out_scores[q * k + j] = res[j].score;
            }
            if (out_counts) out_counts[q] = (uint32_t)res.size();
        }
        todo.swap(again);
        k_sel = std::min(n, k_sel * 4);
    }
    if (timed_out(tctx)) {
        for (size_t i = 0; i < nq * k; i++) {
            if (out_labels) out_labels[i] = (size_t)-1;
            if (out_scores) out_scores[i] = nan;
        }
        if (out_counts)
            for (size_t q = 0; q < nq; q++) out_counts[q] = 0;
        return 1;
    }
    return 0;
}

VecSimQueryReply *FlatMultiIndex::topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) {
    auto *rep = new VecSimQueryReply();
    const size_t cap = std::min(k, indexLabelCount());
    if (cap == 0) {
        last_mode_ = STANDARD_KNN;
        if (k && timed_out(qp ? qp->timeoutCtx : nullptr)) rep->code = VecSim_QueryReply_TimedOut;
        return rep;
    }
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

// unique_results_container: each label once, with its minimum score (brute_force.h:304-321)
VecSimQueryReply *FlatMultiIndex::rangeQuery(const void *blob, double radius, VecSimQueryParams *qp, VecSimQueryReply_Order order) {
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
    std::vector<uint8_t> q(stored_size_);
    preprocess(blob, q.data());
    const double r = type_ == VecSimType_FLOAT64 ? radius : (double)(float)radius;
    size_t cap = 4096, count = 0;
    std::vector<uint64_t> lab;
    std::vector<double> sc;
    for (int attempt = 0; attempt < 2; attempt++) {
        lab.resize(cap);
        sc.resize(cap);
        const int rc = vsgpu_range(store_, q.data(), r, cap, lab.data(), sc.data(), nullptr, &count);
        if (rc == VSGPU_OK) break;
        if (rc != VSGPU_ERR_OVERFLOW) return rep;
        cap = count;
    }
    if (timed_out(tctx)) {
        rep->code = VecSim_QueryReply_TimedOut;
        return rep;
    }
    std::unordered_map<size_t, double> best;
    best.reserve(count * 2);
    for (size_t i = 0; i < count; i++) {
        auto ins = best.emplace((size_t)lab[i], sc[i]);
        if (!ins.second && sc[i] < ins.first->second) ins.first->second = sc[i];
    }
    rep->results.reserve(best.size());
    for (auto &p : best) rep->results.push_back({p.first, p.second});
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

// BFM_BatchIterator::calculateScores (bfm_batch_iterator.h:24-53): one (min score, label) pair per label
int FlatMultiIndex::allScores(const void *processed_query, std::vector<std::pair<double, size_t>> &out) {
    std::lock_guard<std::mutex> g(mu_);
    out.clear();
    if (flush() != 0) return -1;
    const size_t n = id_to_label_.size();
    if (n == 0) return 0;
    std::vector<double> sc(n);
    if (vsgpu_scores(store_, processed_query, sc.data()) != VSGPU_OK) return -1;
    std::unordered_map<size_t, double> best;
    best.reserve(label_to_ids_.size() * 2);
    for (size_t i = 0; i < n; i++) {
        auto ins = best.emplace(id_to_label_[i], sc[i]);
        if (!ins.second && ins.first->second > sc[i]) ins.first->second = sc[i];
    }
    out.reserve(best.size());
    for (auto &p : best) out.emplace_back(p.second, p.first);
    return 0;
}

namespace {
class FlatMultiBatchIterator final : public VecSimBatchIterator {
  public:
    FlatMultiBatchIterator(FlatMultiIndex *idx, std::vector<uint8_t> q, void *tctx)
        : idx_(idx), query_(std::move(q)), tctx_(tctx), label_count_(idx->indexLabelCount()) {}
    VecSimQueryReply *next(size_t n, VecSimQueryReply_Order order) override {
        auto *rep = new VecSimQueryReply();
        if (!computed_) {
            if (idx_->allScores(query_.data(), scores_) != 0) return rep;
            label_count_ = scores_.size();
            computed_ = true;
        }
        if (timed_out(tctx_)) {
            rep->code = VecSim_QueryReply_TimedOut;
            return rep;
        }
        const size_t remaining = scores_.size() - pos_;
        n = std::min(n, remaining);
        auto b = scores_.begin() + (ptrdiff_t)pos_;
        if (n < remaining) std::nth_element(b, b + (ptrdiff_t)n, scores_.end());
        std::sort(b, b + (ptrdiff_t)n);
        rep->results.resize(n);
        for (size_t i = 0; i < n; i++) rep->results[i] = {scores_[pos_ + i].second, scores_[pos_ + i].first};
        pos_ += n;
        returned_ += n;
        if (order == BY_ID)
            std::sort(rep->results.begin(), rep->results.end(),
                      [](const VecSimQueryResult &a, const VecSimQueryResult &c) { return a.id < c.id; });
        return rep;
    }
    bool hasNext() override { return returned_ != label_count_; }
    void reset() override {
        scores_.clear();
        computed_ = false;
        pos_ = returned_ = 0;
    }

  private:
    FlatMultiIndex *idx_;
    std::vector<uint8_t> query_;
    void *tctx_;
    size_t label_count_;
    std::vector<std::pair<double, size_t>> scores_;
    bool computed_ = false;
    size_t pos_ = 0, returned_ = 0;
};
} // namespace

VecSimBatchIterator *FlatMultiIndex::newBatchIterator(const void *blob, VecSimQueryParams *qp) {
    return new FlatMultiBatchIterator(this, preprocessQuery(blob), qp ? qp->timeoutCtx : nullptr);
}

VecSimIndexBasicInfo FlatMultiIndex::basicInfo() {
    VecSimIndexBasicInfo b{};
    b.algo = VecSimAlgo_BF;
    b.metric = metric_;
    b.type = type_;
    b.isMulti = true;
    b.isTiered = false;
    b.isDisk = false;
    b.blockSize = block_size_;
    b.dim = dim_;
    return b;
}

VecSimIndexStatsInfo FlatMultiIndex::statsInfo() {
    VecSimIndexStatsInfo s{};
    s.memory = sizeof(*this) + id_to_label_.capacity() * sizeof(size_t) + label_to_ids_.size() * 48 + id_to_label_.size() * sizeof(idType) +
               pending_rows_.capacity() + vsgpu_store_device_bytes(store_);
    return s;
}

VecSimIndexDebugInfo FlatMultiIndex::debugInfo() {
    VecSimIndexDebugInfo d{};
    d.commonInfo.basicInfo = basicInfo();
    d.commonInfo.indexSize = indexSize();
    d.commonInfo.indexLabelCount = indexLabelCount();
    d.commonInfo.memory = statsInfo().memory;
    d.commonInfo.lastMode = last_mode_;
    return d;
}

bool FlatMultiIndex::preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) {
    // subsetSize counts labels; every vector of a label is scored ad hoc, so scale by vectors per label
    const size_t labels = std::max<size_t>(indexLabelCount(), 1);
    const size_t rows = std::min(indexSize(), (size_t)((double)std::min(subsetSize, labels) * (double)indexSize() / (double)labels));
    const bool res = prefer_adhoc_flat(indexSize(), stored_size_, rows, k);
    last_mode_ = res ? (initial_check ? HYBRID_ADHOC_BF : HYBRID_BATCHES_TO_ADHOC_BF) : HYBRID_BATCHES;
    return res;
}

} // namespace vsb
