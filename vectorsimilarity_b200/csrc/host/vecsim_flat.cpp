// Flat (brute-force) index: host bookkeeping of the reference's BruteForceIndex_Single
// (/root/reference/src/VecSim/algorithms/brute_force/brute_force.h:174-326, brute_force_single.h:134-212,
// bf_batch_iterator.h:59-214) over a device-resident row store. All distance work and selection runs in
// libvsgpu.so; there is no CPU fallback — construction fails when no device is available.
#include "vecsim_index.h"
#include "vecsim_numeric.h"
#include "vecsim_hybrid.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <functional>
#include <limits>
#include <stdexcept>

namespace vsb {

Globals &globals() {
    static Globals g;
    return g;
}

namespace {
struct alignas(16) HookHeader {
    size_t size;
    void (*free_fn)(void *);
};
std::atomic<size_t> g_hook_bytes{0};
} // namespace

void *hook_alloc(size_t n) {
    const Globals &g = globals();
    void *(*alloc_fn)(size_t) = (g.mem_set && g.mem.allocFunction) ? g.mem.allocFunction : std::malloc;
    void (*free_fn)(void *) = (g.mem_set && g.mem.allocFunction && g.mem.freeFunction) ? g.mem.freeFunction : std::free;
    auto *h = static_cast<HookHeader *>(alloc_fn(n + sizeof(HookHeader)));
    if (!h) throw std::bad_alloc();
    h->size = n;
    h->free_fn = free_fn;
    g_hook_bytes.fetch_add(n + sizeof(HookHeader), std::memory_order_relaxed);
    return h + 1;
}

void hook_free(void *p) noexcept {
    if (!p) return;
    HookHeader *h = static_cast<HookHeader *>(p) - 1;
    g_hook_bytes.fetch_sub(h->size + sizeof(HookHeader), std::memory_order_relaxed);
    h->free_fn(h);
}

size_t hook_bytes() { return g_hook_bytes.load(std::memory_order_relaxed); }

size_t type_size(VecSimType t) {
    switch (t) {
    case VecSimType_FLOAT32: return 4;
    case VecSimType_FLOAT64: return 8;
    case VecSimType_BFLOAT16:
    case VecSimType_FLOAT16: return 2;
    case VecSimType_INT8:
    case VecSimType_UINT8: return 1;
    case VecSimType_INT32: return 4;
    case VecSimType_INT64: return 8;
    }
    return 0;
}

size_t stored_size(VecSimType t, size_t dim, VecSimMetric m) {
    size_t s = type_size(t) * dim;
    if (m == VecSimMetric_Cosine && (t == VecSimType_INT8 || t == VecSimType_UINT8)) s += sizeof(float);
    return s;
}

void normalize_blob(void *blob, size_t dim, VecSimType type) {
    switch (type) {
    case VecSimType_FLOAT32: normalize_f32((float *)blob, dim); break;
    case VecSimType_FLOAT64: normalize_f64((double *)blob, dim); break;
    case VecSimType_BFLOAT16: normalize_16<true>((uint16_t *)blob, dim); break;
    case VecSimType_FLOAT16: normalize_16<false>((uint16_t *)blob, dim); break;
    case VecSimType_INT8: append_int_norm<int8_t>(blob, dim); break;
    case VecSimType_UINT8: append_int_norm<uint8_t>(blob, dim); break;
    default: break;
    }
}

static constexpr size_t FLUSH_ROWS = 8192;
static constexpr size_t FLUSH_BYTES = (size_t)64 << 20;

FlatIndex::FlatIndex(const BFParams &p, void *logCtx, int device)
    : type_(p.type), metric_(p.metric), dim_(p.dim), block_size_(p.blockSize ? p.blockSize : 1024),
      data_size_(type_size(p.type) * p.dim), stored_size_(stored_size(p.type, p.dim, p.metric)), log_ctx_(logCtx) {
    store_ = vsgpu_store_create(device >= 0 ? device : globals().device, (int)type_, (int)metric_, dim_, p.initialCapacity);
}

FlatIndex::~FlatIndex() {
    if (store_) vsgpu_store_destroy(store_);
}

// preprocessForStorage / preprocessQuery (spaces/computer/preprocessors.h:49-141): cosine only.
void FlatIndex::preprocess(const void *blob, uint8_t *out) const {
    std::memcpy(out, blob, data_size_);
    if (metric_ == VecSimMetric_Cosine) normalize_blob(out, dim_, type_);
}

std::vector<uint8_t> FlatIndex::preprocessQuery(const void *blob) {
    std::vector<uint8_t> q(stored_size_);
    preprocess(blob, q.data());
    return q;
}

int FlatIndex::flush() {
    if (pending_labels_.empty()) return 0;
    int rc = vsgpu_store_append(store_, pending_rows_.data(), stored_size_, pending_labels_.data(), pending_labels_.size());
    if (rc != VSGPU_OK) return rc;
    pending_rows_.clear();
    pending_labels_.clear();
    return 0;
}

vsgpu_store *FlatIndex::deviceStore() {
    std::lock_guard<std::mutex> g(mu_);
    if (flush() != 0) return nullptr;
    return store_;
}

void FlatIndex::lastStats(vsgpu_stats *out) { vsgpu_last_stats(store_, out); }

bool FlatIndex::findId(size_t label, idType *id) const {
    if (identity_) {
        if (label >= count_) return false;
        *id = (idType)label;
        return true;
    }
    auto it = label_to_id_.find(label);
    if (it == label_to_id_.end()) return false;
    *id = it->second;
    return true;
}

void FlatIndex::materialize() {
    if (!identity_) return;
    id_to_label_.resize(count_);
    label_to_id_.reserve(count_ * 2 + 16);
    for (size_t i = 0; i < count_; i++) {
        id_to_label_[i] = i;
        label_to_id_.emplace(i, (idType)i);
    }
    identity_ = false;
}

long FlatIndex::appendDeviceRows(const void *dev_rows, size_t stride, size_t n, size_t first_label) {
    std::lock_guard<std::mutex> g(mu_);
    if (n == 0) return 0;
    if (flush() != 0) return -1;
    if (count_ + n >= 0xfffffffeull) return -1;
    for (size_t i = 0; i < n && !(identity_ && first_label >= count_); i++) {
        idType tmp;
        if (findId(first_label + i, &tmp)) return -1; // bulk ingest never overwrites
    }
    if (vsgpu_store_append_device(store_, dev_rows, stride, nullptr, first_label, nullptr, n) != VSGPU_OK) return -1;
    if (count_ > 0 && first_label <= max_label_) labels_monotone_ = false;
    if (!(identity_ && first_label == count_)) {
        materialize();
        for (size_t i = 0; i < n; i++) {
            id_to_label_.push_back(first_label + i);
            label_to_id_.emplace(first_label + i, (idType)(count_ + i));
        }
    }
    max_label_ = count_ == 0 ? first_label + n - 1 : std::max(max_label_, first_label + n - 1);
    count_ += n;
    return (long)n;
}

int FlatIndex::addVector(const void *blob, size_t label) {
    std::lock_guard<std::mutex> g(mu_);
    idType existing;
    if (findId(label, &existing)) {
        // Overwrite in place. The reference copies the caller's raw bytes here
        // (brute_force_single.h:138-144), skipping cosine preprocessing; we preprocess.
        if (flush() != 0) return -1;
        std::vector<uint8_t> row(stored_size_);
        preprocess(blob, row.data());
        if (vsgpu_store_update(store_, existing, row.data(), label) != VSGPU_OK) return -1;
        return 0;
    }
    const size_t id = count_;
    if (id >= 0xfffffffeull) return -1;
    if (id > 0 && label <= max_label_) labels_monotone_ = false;
    max_label_ = id == 0 ? label : std::max(max_label_, label);
    if (!(identity_ && label == id)) {
        materialize();
        id_to_label_.push_back(label);
        label_to_id_.emplace(label, (idType)id);
    }
    count_++;
    const size_t off = pending_rows_.size();
    pending_rows_.resize(off + stored_size_);
    preprocess(blob, pending_rows_.data() + off);
    pending_labels_.push_back(label);
    if (pending_labels_.size() >= FLUSH_ROWS || pending_rows_.size() >= FLUSH_BYTES)
        if (flush() != 0) return -1;
    return 1;
}

long FlatIndex::addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) {
    long added = 0;
    for (size_t i = 0; i < n; i++) {
        const int rc = addVector((const uint8_t *)blobs + i * data_size_, labels ? labels[i] : first_label + i);
        if (rc < 0) return -1;
        added += rc;
    }
    return added;
}

int FlatIndex::deleteVector(size_t label) {
    std::lock_guard<std::mutex> g(mu_);
    idType id;
    if (!findId(label, &id)) return 0;
    if (flush() != 0) return 0;
    const size_t last = count_ - 1;
    if (vsgpu_store_remove_swap(store_, id) != VSGPU_OK) return 0;
    if (identity_ && id == last) {
        count_--;
        if (count_ == 0) max_label_ = 0;
        return 1;
    }
    materialize();
    label_to_id_.erase(label);
    if (id != last) {
        // the last row moved into the hole (brute_force.h:204-218)
        const size_t moved = id_to_label_[last];
        id_to_label_[id] = moved;
        label_to_id_[moved] = id;
        labels_monotone_ = false;
    }
    id_to_label_.pop_back();
    count_--;
    if (count_ == 0) {
        labels_monotone_ = true;
        identity_ = true;
        label_to_id_.clear();
        max_label_ = 0;
    }
    return 1;
}

double FlatIndex::getDistanceFrom(size_t label, const void *blob) {
    std::lock_guard<std::mutex> g(mu_);
    idType found;
    if (!findId(label, &found)) return std::numeric_limits<double>::quiet_NaN();
    if (flush() != 0) return std::numeric_limits<double>::quiet_NaN();
    const uint32_t id = found;
    double out = std::numeric_limits<double>::quiet_NaN();
    // the caller's blob is used as is (brute_force_single.h:200-212)
    if (vsgpu_distances(store_, blob, &id, 1, &out) != VSGPU_OK) return std::numeric_limits<double>::quiet_NaN();
    return out;
}

void FlatIndex::exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) {
    std::lock_guard<std::mutex> g(mu_);
    const double nan = std::numeric_limits<double>::quiet_NaN();
    for (size_t i = 0; i < n; i++) out[i] = nan;
    if (flush() != 0) return;
    std::vector<uint32_t> ids;
    std::vector<size_t> pos;
    for (size_t i = 0; i < n; i++) {
        idType id;
        if (!findId(labels[i], &id)) continue;
        ids.push_back(id);
        pos.push_back(i);
    }
    if (ids.empty()) return;
    std::vector<double> d(ids.size());
    if (vsgpu_distances(store_, processed_query, ids.data(), ids.size(), d.data()) != VSGPU_OK) return;
    for (size_t j = 0; j < ids.size(); j++) out[pos[j]] = d[j];
}

// SURVEY App. A2. `ids`-ordered candidates: ascending (score, internal id), cnt >= min(2k, n).
void FlatIndex::resolve(const size_t *labels, const double *scores, const uint32_t *ids, size_t cnt, size_t k,
                        std::vector<VecSimQueryResult> &out) {
    out.clear();
    if (cnt <= k) {
        for (size_t i = 0; i < cnt; i++) out.push_back({labels[i], scores[i]});
    } else {
        const double T = scores[k - 1];
        size_t c = 0;
        while (c < cnt && scores[c] < T) c++;
        size_t e = c;
        while (e < cnt && scores[e] == T) e++;
        // P: the first k rows in scan (internal id) order among score <= T
        std::vector<uint32_t> le(ids, ids + e);
        std::nth_element(le.begin(), le.begin() + (k - 1), le.end());
        const uint32_t id_cut = le[k - 1];
        std::vector<size_t> tie_labels;
        for (size_t i = c; i < e; i++)
            if (ids[i] <= id_cut) tie_labels.push_back(labels[i]);
        std::sort(tie_labels.begin(), tie_labels.end());
        const size_t m = k - c;
        for (size_t i = 0; i < c; i++) out.push_back({labels[i], scores[i]});
        for (size_t i = 0; i < m && i < tie_labels.size(); i++) out.push_back({tie_labels[i], T});
    }
    // reply order: ascending (score, label) — brute_force.h:284-288 drains a max-heap of pairs
    std::sort(out.begin(), out.end(), [](const VecSimQueryResult &a, const VecSimQueryResult &b) {
        if (a.score < b.score) return true;
        if (b.score < a.score) return false;
        return a.id < b.id;
    });
}

int FlatIndex::topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *out_labels,
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
    if (k == 0 || count_ == 0) {
        pad_all();
        return 0;
    }
    if (timed_out(tctx)) {
        pad_all();
        return 1;
    }
    if (flush() != 0) return -1;
    const size_t n = count_;
    // L2 / IP queries need no preprocessing: hand the caller's buffer to the device layer as it is (it stages it in pinned
    // memory once) instead of copying 3 MB of batch through a second host buffer first
    const bool as_is = metric_ != VecSimMetric_Cosine && stored_size_ == data_size_;
    std::vector<uint8_t> qbuf;
    if (!as_is) {
        qbuf.resize(nq * stored_size_);
        for (size_t q = 0; q < nq; q++) preprocess((const uint8_t *)queries + q * data_size_, qbuf.data() + q * stored_size_);
    }
    const void *q_src = as_is ? queries : (const void *)qbuf.data();
    const bool fast = labels_monotone_;
    size_t k_sel = std::min(k, n);
    if (!fast) k_sel = (k > n / 2) ? n : std::min(2 * k, n);
    std::vector<uint64_t> lab;
    std::vector<double> sc;
    std::vector<uint32_t> ids;
    uint64_t *lab_p;
    double *sc_p;
    static_assert(sizeof(size_t) == sizeof(uint64_t), "labelType is 64-bit");
    if (fast && k_sel == k && out_labels && out_scores) {
        lab_p = (uint64_t *)out_labels;
        sc_p = out_scores;
    } else {
        lab.resize(nq * k_sel);
        sc.resize(nq * k_sel);
        lab_p = lab.data();
        sc_p = sc.data();
    }
    if (!fast) ids.resize(nq * k_sel);
    const int rc = vsgpu_topk(store_, q_src, nq, stored_size_, k_sel, (unsigned)globals().topk_mode, lab_p, sc_p,
                              fast ? nullptr : ids.data(), nullptr);
    if (rc != VSGPU_OK) return -1;
    if (timed_out(tctx)) {
        pad_all();
        return 1;
    }
    if (fast) {
        // labels grow with the internal id: (score, id) order is (score, label) order
        if (lab_p != (uint64_t *)out_labels) {
            for (size_t q = 0; q < nq; q++)
                for (size_t j = 0; j < k; j++) {
                    const bool valid = j < k_sel;
                    if (out_labels) out_labels[q * k + j] = valid ? lab[q * k_sel + j] : (size_t)-1;
                    if (out_scores) out_scores[q * k + j] = valid ? sc[q * k_sel + j] : nan;
                }
        }
        if (out_counts)
            for (size_t q = 0; q < nq; q++) out_counts[q] = (uint32_t)k_sel;
        return 0;
    }
    std::vector<VecSimQueryResult> res;
    for (size_t q = 0; q < nq; q++) {
        resolve((const size_t *)lab.data() + q * k_sel, sc.data() + q * k_sel, ids.data() + q * k_sel, k_sel, k, res);
        for (size_t j = 0; j < k; j++) {
            const bool valid = j < res.size();
            if (out_labels) out_labels[q * k + j] = valid ? res[j].id : (size_t)-1;
            if (out_scores) out_scores[q * k + j] = valid ? res[j].score : nan;
        }
        if (out_counts) out_counts[q] = (uint32_t)res.size();
    }
    return 0;
}

VecSimQueryReply *FlatIndex::topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) {
    auto *rep = new VecSimQueryReply();
    if (k == 0) {
        last_mode_ = STANDARD_KNN;
        return rep;
    }
    if (indexSize() == 0) {
        last_mode_ = STANDARD_KNN;
        if (timed_out(qp ? qp->timeoutCtx : nullptr)) rep->code = VecSim_QueryReply_TimedOut;
        return rep;
    }
    auto run_one = [&](const void *b, size_t kk, VecSimQueryParams *p, VecSimQueryReply *out) {
        const size_t cap = std::min(kk, indexSize());
        if (cap == 0) return;
        std::vector<size_t> labels(cap);
        std::vector<double> scores(cap);
        uint32_t cnt = 0;
        const int rc = topKBatch(b, 1, cap, p, labels.data(), scores.data(), &cnt);
        if (rc == 1) out->code = VecSim_QueryReply_TimedOut;
        if (rc != 0) return;
        out->results.resize(cnt);
        for (uint32_t i = 0; i < cnt; i++) out->results[i] = {labels[i], scores[i]};
    };
    if (qp && qp->timeoutCtx) { // a caller with its own timeout context is served on its own
        run_one(blob, k, qp, rep);
        return rep;
    }
    // Flat combining: whoever finds no leader becomes one and serves every query that is waiting — its own included — as ONE
    // batched device call; callers that arrive while that call runs wait and are served by the next round. A lone caller pays
    // nothing; N concurrent callers share one pass over the store instead of queueing N passes behind the index mutex.
    PendingQuery me{blob, k, rep};
    std::unique_lock<std::mutex> ql(q_mu_);
    q_wait_.push_back(&me);
    if (q_leader_) {
        q_cv_.wait(ql, [&] { return me.done || !q_leader_; });
        if (me.done) return rep;
    }
    q_leader_ = true;
    while (!q_wait_.empty()) {
        std::vector<PendingQuery *> batch;
        batch.swap(q_wait_);
        ql.unlock();
        if (batch.size() == 1) {
            run_one(batch[0]->blob, batch[0]->k, nullptr, batch[0]->rep);
        } else {
            size_t kmax = 0;
            for (PendingQuery *p : batch) kmax = std::max(kmax, p->k);
            kmax = std::min(kmax, indexSize());
            const size_t nq = batch.size();
            std::vector<uint8_t> qbuf(nq * data_size_);
            for (size_t i = 0; i < nq; i++) std::memcpy(qbuf.data() + i * data_size_, batch[i]->blob, data_size_);
            std::vector<size_t> labels(nq * std::max<size_t>(kmax, 1));
            std::vector<double> scores(nq * std::max<size_t>(kmax, 1));
            std::vector<uint32_t> cnt(nq, 0);
            const int rc = kmax ? topKBatch(qbuf.data(), nq, kmax, nullptr, labels.data(), scores.data(), cnt.data()) : 0;
            for (size_t i = 0; i < nq; i++) {
                if (rc != 0) continue;
                // the k best of a query are a prefix of its kmax best (same total order)
                const size_t take = std::min<size_t>(std::min<size_t>(cnt[i], batch[i]->k), kmax);
                batch[i]->rep->results.resize(take);
                for (size_t j = 0; j < take; j++) batch[i]->rep->results[j] = {labels[i * kmax + j], scores[i * kmax + j]};
            }
        }
        ql.lock();
        combined_batches_++;
        combined_queries_ += batch.size();
        for (PendingQuery *p : batch) p->done = true;
        q_cv_.notify_all();
    }
    q_leader_ = false;
    q_cv_.notify_all(); // a caller that slipped in after the last swap takes over as leader
    return rep;
}

VecSimQueryReply *FlatIndex::rangeQuery(const void *blob, double radius, VecSimQueryParams *qp,
                                        VecSimQueryReply_Order order) {
    std::lock_guard<std::mutex> g(mu_);
    auto *rep = new VecSimQueryReply();
    void *tctx = qp ? qp->timeoutCtx : nullptr;
    last_mode_ = RANGE_QUERY;
    if (count_ == 0) return rep;
    if (timed_out(tctx)) {
        rep->code = VecSim_QueryReply_TimedOut; // partial (here: empty) results, brute_force.h:311-314
        return rep;
    }
    if (flush() != 0) return rep;
    std::vector<uint8_t> q(stored_size_);
    preprocess(blob, q.data());
    // the bound is compared in the index's DistType (brute_force.h:306)
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
    rep->results.resize(count);
    for (size_t i = 0; i < count; i++) rep->results[i] = {(size_t)lab[i], sc[i]};
    auto by_id = [](const VecSimQueryResult &a, const VecSimQueryResult &b) { return a.id < b.id; };
    auto by_score_id = [](const VecSimQueryResult &a, const VecSimQueryResult &b) {
        if (a.score < b.score) return true;
        if (b.score < a.score) return false;
        return a.id < b.id;
    };
    if (order == BY_ID) std::sort(rep->results.begin(), rep->results.end(), by_id);
    else std::sort(rep->results.begin(), rep->results.end(), by_score_id);
    return rep;
}

int FlatIndex::allScores(const void *processed_query, std::vector<std::pair<double, size_t>> &out) {
    std::lock_guard<std::mutex> g(mu_);
    out.clear();
    if (flush() != 0) return -1;
    const size_t n = count_;
    if (n == 0) return 0;
    std::vector<double> sc(n);
    if (vsgpu_scores(store_, processed_query, sc.data()) != VSGPU_OK) return -1;
    out.resize(n);
    for (size_t i = 0; i < n; i++) out[i] = {sc[i], labelOf(i)};
    return 0;
}

// bf_batch_iterator.h:59-214 — first call scores every row (on the device), later calls only select.
class FlatBatchIterator final : public VecSimBatchIterator {
  public:
    using ScoreFn = std::function<int(const void *, std::vector<std::pair<double, size_t>> &)>;
    FlatBatchIterator(ScoreFn score_all, size_t label_count, std::vector<uint8_t> q, void *tctx)
        : score_all_(std::move(score_all)), query_(std::move(q)), tctx_(tctx), label_count_(label_count) {}
    VecSimQueryReply *next(size_t n, VecSimQueryReply_Order order) override {
        auto *rep = new VecSimQueryReply();
        if (!computed_) {
            if (score_all_(query_.data(), scores_) != 0) return rep;
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
    ScoreFn score_all_;
    std::vector<uint8_t> query_;
    void *tctx_;
    size_t label_count_;
    std::vector<std::pair<double, size_t>> scores_;
    bool computed_ = false;
    size_t pos_ = 0, returned_ = 0;
};

VecSimBatchIterator *new_flat_batch_iterator(std::function<int(const void *, std::vector<std::pair<double, size_t>> &)> score_all,
                                             size_t label_count, std::vector<uint8_t> query, void *tctx) {
    return new FlatBatchIterator(std::move(score_all), label_count, std::move(query), tctx);
}

VecSimBatchIterator *FlatIndex::newBatchIterator(const void *blob, VecSimQueryParams *qp) {
    return new_flat_batch_iterator([this](const void *q, std::vector<std::pair<double, size_t>> &out) { return allScores(q, out); },
                                   indexLabelCount(), preprocessQuery(blob), qp ? qp->timeoutCtx : nullptr);
}

VecSimIndexBasicInfo FlatIndex::basicInfo() {
    VecSimIndexBasicInfo b{};
    b.algo = VecSimAlgo_BF;
    b.metric = metric_;
    b.type = type_;
    b.isMulti = false;
    b.isTiered = false;
    b.isDisk = false;
    b.blockSize = block_size_;
    b.dim = dim_;
    return b;
}

VecSimIndexStatsInfo FlatIndex::statsInfo() {
    VecSimIndexStatsInfo s{};
    s.memory = sizeof(*this) + id_to_label_.capacity() * sizeof(size_t) + label_to_id_.size() * 32 +
               pending_rows_.capacity() + vsgpu_store_device_bytes(store_);
    return s;
}

VecSimIndexDebugInfo FlatIndex::debugInfo() {
    VecSimIndexDebugInfo d{};
    d.commonInfo.basicInfo = basicInfo();
    d.commonInfo.indexSize = indexSize();
    d.commonInfo.indexLabelCount = indexLabelCount();
    d.commonInfo.memory = statsInfo().memory;
    d.commonInfo.lastMode = last_mode_;
    return d;
}

// The reference's tree (brute_force.h:380-451) is fitted to CPU costs; here the decision is a cost comparison with device
// rates (vecsim_hybrid.h).
bool FlatIndex::preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) {
    const bool res = prefer_adhoc_flat(indexSize(), stored_size_, subsetSize, k);
    last_mode_ = res ? (initial_check ? HYBRID_ADHOC_BF : HYBRID_BATCHES_TO_ADHOC_BF) : HYBRID_BATCHES;
    return res;
}

} // namespace vsb

