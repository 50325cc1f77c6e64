// Flat index with its rows sharded over several devices of ONE process (SURVEY.md §5 "Distributed communication backend",
// §8b "Additions", §8e): VecSimGPU_Configure(devices) makes VecSimIndex_New(VecSimAlgo_BF) return this class, so a program
// that links libvecsim_b200.so — RediSearch — gets all the GPUs of the box through the unchanged VecSimIndex_* API.
// The reference has no counterpart: its flat index is one in-process scan (algorithms/brute_force/brute_force.h:242-291).
//
// One FlatIndex (host bookkeeping + device store) per device and one persistent host thread per shard, so the ~20 kernel
// launches of a shard's scan are enqueued on all devices at the same time. A batched top-k runs through vsgpu_group_*
// (include/vsgpu.h): per shard H2D + scan + pack + peer copy of the packed top-k into the root device, device merge on the
// root, one host wait at the end. Range queries and batch iterators concatenate the shards' replies on the host.
//
// Routing: a label lives on shard (label mod shards), except labels inside a range that was bulk-ingested from device
// memory (VecSimGPU_AppendDeviceRows), which live where their rows were. Reply order is the reference's, ascending
// (score, label). Ties AT the k-th score: the reference admits them in scan (insertion) order; here each shard does, and the
// shards' lists are merged by label — identical whenever labels grow with insertion order (bulk ingest, label = id), the only
// case in which the single-device index skips its host-side tie resolution too. Otherwise the per-shard replies are resolved
// on the host exactly like the single-device index and merged there.
#include "vecsim_index.h"
#include "vecsim_hybrid.h"
#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <limits>
#include <thread>

namespace vsb {

// One persistent thread per shard: run(fn) executes fn(shard) on every worker and waits for all of them.
class ShardWorkers {
  public:
    explicit ShardWorkers(size_t n) : tasks_(n), state_(n, 0) {
        for (size_t i = 0; i < n; i++) threads_.emplace_back([this, i] { loop(i); });
    }
    ~ShardWorkers() {
        {
            std::lock_guard<std::mutex> g(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
    }
    void run(const std::function<void(size_t)> &fn) {
        {
            std::lock_guard<std::mutex> g(mu_);
            for (size_t i = 0; i < tasks_.size(); i++) {
                tasks_[i] = &fn;
                state_[i] = 1;
            }
            pending_ = tasks_.size();
        }
        cv_.notify_all();
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [this] { return pending_ == 0; });
    }

  private:
    void loop(size_t i) {
        for (;;) {
            const std::function<void(size_t)> *fn;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_.wait(g, [&] { return stop_ || state_[i] == 1; });
                if (stop_) return;
                fn = tasks_[i];
                state_[i] = 2;
            }
            (*fn)(i);
            {
                std::lock_guard<std::mutex> g(mu_);
                state_[i] = 0;
                if (--pending_ == 0) done_cv_.notify_all();
            }
        }
    }
    std::vector<std::thread> threads_;
    std::vector<const std::function<void(size_t)> *> tasks_;
    std::vector<int> state_;
    size_t pending_ = 0;
    bool stop_ = false;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
};

ShardedFlatIndex::ShardedFlatIndex(const BFParams &p, void *logCtx, const std::vector<int> &devices)
    : devices_(devices), params_(p), data_size_(type_size(p.type) * p.dim), stored_size_(stored_size(p.type, p.dim, p.metric)) {
    if (devices.empty() || p.type == VecSimType_FLOAT64) return; // fp64 scores do not fit the packed hit: not sharded
    BFParams sp = p;
    sp.initialCapacity = (p.initialCapacity + devices.size() - 1) / devices.size();
    std::vector<vsgpu_store *> stores;
    for (int d : devices) {
        auto f = std::make_unique<FlatIndex>(sp, logCtx, d);
        if (!f->ok()) return;
        stores.push_back(f->deviceStore());
        shards_.push_back(std::move(f));
    }
    group_ = vsgpu_group_create(stores.data(), stores.size());
    if (group_) workers_ = std::make_unique<ShardWorkers>(shards_.size());
}

ShardedFlatIndex::~ShardedFlatIndex() {
    workers_.reset();
    if (group_) vsgpu_group_destroy(group_);
    shards_.clear();
}

size_t ShardedFlatIndex::route(size_t label) const {
    // ranges_ is sorted by first label and its entries do not overlap
    auto it = std::upper_bound(ranges_.begin(), ranges_.end(), label, [](size_t l, const Range &r) { return l < r.first; });
    if (it != ranges_.begin()) {
        --it;
        if (label <= it->last) return it->shard;
    }
    return label % shards_.size();
}

int ShardedFlatIndex::addVector(const void *blob, size_t label) { return shards_[route(label)]->addVector(blob, label); }

long ShardedFlatIndex::addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) {
    long added = 0;
    for (size_t i = 0; i < n; i++) {
        const int rc = addVector((const uint8_t *)blobs + i * data_size_, labels ? labels[i] : first_label + i);
        if (rc < 0) return -1;
        added += rc;
    }
    return added;
}

long ShardedFlatIndex::appendDeviceRows(const void *dev_rows, size_t stride, size_t n, size_t first_label) {
    std::lock_guard<std::mutex> g(mu_);
    if (n == 0) return 0;
    const int dev = vsgpu_pointer_device(dev_rows);
    size_t sh = shards_.size();
    for (size_t i = 0; i < devices_.size(); i++)
        if (devices_[i] == dev) sh = i;
    if (sh == shards_.size()) return -1; // the rows must already be on one of the shards' devices
    // the new range must not overlap an earlier one, and no label of it may already live elsewhere by the modulo rule
    for (const Range &r : ranges_)
        if (first_label <= r.last && r.first <= first_label + n - 1) return -1;
    const long rc = shards_[sh]->appendDeviceRows(dev_rows, stride, n, first_label);
    if (rc < 0) return rc;
    ranges_.push_back({first_label, first_label + n - 1, sh});
    std::sort(ranges_.begin(), ranges_.end(), [](const Range &a, const Range &b) { return a.first < b.first; });
    return rc;
}

int ShardedFlatIndex::deleteVector(size_t label) { return shards_[route(label)]->deleteVector(label); }

double ShardedFlatIndex::getDistanceFrom(size_t label, const void *blob) { return shards_[route(label)]->getDistanceFrom(label, blob); }

size_t ShardedFlatIndex::indexSize() {
    size_t n = 0;
    for (auto &s : shards_) n += s->indexSize();
    return n;
}

std::vector<uint8_t> ShardedFlatIndex::preprocessQuery(const void *blob) { return shards_[0]->preprocessQuery(blob); }

void ShardedFlatIndex::exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) {
    std::vector<std::vector<size_t>> lab(shards_.size()), pos(shards_.size());
    for (size_t i = 0; i < n; i++) {
        const size_t sh = route(labels[i]);
        lab[sh].push_back(labels[i]);
        pos[sh].push_back(i);
    }
    workers_->run([&](size_t sh) {
        if (lab[sh].empty()) return;
        std::vector<double> d(lab[sh].size());
        shards_[sh]->exactDistances(processed_query, lab[sh].data(), d.data(), d.size());
        for (size_t j = 0; j < d.size(); j++) out[pos[sh][j]] = d[j];
    });
}

int ShardedFlatIndex::topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *out_labels,
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
    const size_t total = indexSize();
    if (k == 0 || total == 0) {
        pad_all();
        return 0;
    }
    if (timed_out(tctx)) {
        pad_all();
        return 1;
    }
    bool fast = true;
    for (auto &s : shards_) {
        if (!s->deviceStore()) return -1; // flushes staged rows
        fast = fast && s->labelsMonotone();
    }
    static_assert(sizeof(size_t) == sizeof(uint64_t), "labelType is 64-bit");
    if (fast) {
        // every shard's (score, id) order is its (score, label) order: scan, gather and merge on the devices
        std::vector<uint8_t> qbuf(nq * stored_size_);
        for (size_t q = 0; q < nq; q++) {
            const std::vector<uint8_t> one = shards_[0]->preprocessQuery((const uint8_t *)queries + q * data_size_);
            std::memcpy(qbuf.data() + q * stored_size_, one.data(), stored_size_);
        }
        if (vsgpu_group_topk_begin(group_, qbuf.data(), nq, stored_size_, k) != VSGPU_OK) return -1;
        std::vector<int> rcs(shards_.size(), 0);
        std::vector<std::string> errs(shards_.size());
        const unsigned mode = (unsigned)globals().topk_mode;
        auto failed = [&]() { // error texts are per thread in libvsgpu: carry the first one over to the caller's thread
            for (size_t sh = 0; sh < rcs.size(); sh++)
                if (rcs[sh] != VSGPU_OK) {
                    vsgpu_set_last_error(("shard " + std::to_string(sh) + ": " + errs[sh]).c_str());
                    return true;
                }
            return false;
        };
        workers_->run([&](size_t sh) {
            rcs[sh] = vsgpu_group_topk_shard(group_, sh, nq, k, mode, 0);
            if (rcs[sh] != VSGPU_OK) errs[sh] = vsgpu_last_error();
        });
        if (failed()) return -1;
        std::vector<uint64_t> lab(nq * k);
        std::vector<double> sc(nq * k);
        int rc = vsgpu_group_topk_finish(group_, nq, k, lab.data(), sc.data());
        if (rc == 1) {
            // some shard's candidate buffer overflowed (adversarial ties): the shards redo those queries exactly, push again
            workers_->run([&](size_t sh) {
                rcs[sh] = vsgpu_store_sync(shards_[sh]->deviceStore());
                if (rcs[sh] == VSGPU_OK) rcs[sh] = vsgpu_group_topk_shard(group_, sh, nq, k, mode, 1);
                if (rcs[sh] != VSGPU_OK) errs[sh] = vsgpu_last_error();
            });
            if (failed()) return -1;
            rc = vsgpu_group_topk_finish(group_, nq, k, lab.data(), sc.data());
            if (rc == 1) rc = VSGPU_OK; // flags are sticky for the batch; the lists are exact now
        } else {
            workers_->run([&](size_t sh) { vsgpu_store_sync(shards_[sh]->deviceStore()); }); // completes the shards' stats
        }
        if (rc != VSGPU_OK) return -1;
        last_ms_ = vsgpu_group_last_ms(group_);
        if (timed_out(tctx)) {
            pad_all();
            return 1;
        }
        for (size_t q = 0; q < nq; q++) {
            uint32_t cnt = 0;
            for (size_t j = 0; j < k; j++) {
                const bool valid = lab[q * k + j] != ~0ull;
                cnt += valid;
                if (out_labels) out_labels[q * k + j] = valid ? (size_t)lab[q * k + j] : (size_t)-1;
                if (out_scores) out_scores[q * k + j] = valid ? sc[q * k + j] : nan;
            }
            if (out_counts) out_counts[q] = cnt;
        }
        return 0;
    }
    // general labels: each shard resolves the reference's tie rule on the host (FlatIndex::topKBatch), lists merged here
    const size_t S = shards_.size();
    std::vector<std::vector<size_t>> labs(S, std::vector<size_t>(nq * k));
    std::vector<std::vector<double>> scs(S, std::vector<double>(nq * k));
    std::vector<std::vector<uint32_t>> cnts(S, std::vector<uint32_t>(nq, 0));
    std::vector<int> rcs(S, 0);
    workers_->run([&](size_t sh) {
        if (shards_[sh]->indexSize() == 0) return;
        rcs[sh] = shards_[sh]->topKBatch(queries, nq, k, nullptr, labs[sh].data(), scs[sh].data(), cnts[sh].data());
    });
    for (int rc : rcs)
        if (rc != 0) return -1;
    if (timed_out(tctx)) {
        pad_all();
        return 1;
    }
    std::vector<VecSimQueryResult> all;
    for (size_t q = 0; q < nq; q++) {
        all.clear();
        for (size_t sh = 0; sh < S; sh++)
            for (uint32_t j = 0; j < cnts[sh][q]; j++) all.push_back({labs[sh][q * k + j], scs[sh][q * k + j]});
        std::sort(all.begin(), all.end(), [](const VecSimQueryResult &a, const VecSimQueryResult &b) {
            if (a.score < b.score) return true;
            if (b.score < a.score) return false;
            return a.id < b.id;
        });
        const size_t cnt = std::min(all.size(), k);
        for (size_t j = 0; j < k; j++) {
            if (out_labels) out_labels[q * k + j] = j < cnt ? all[j].id : (size_t)-1;
            if (out_scores) out_scores[q * k + j] = j < cnt ? all[j].score : nan;
        }
        if (out_counts) out_counts[q] = (uint32_t)cnt;
    }
    return 0;
}

VecSimQueryReply *ShardedFlatIndex::topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) {
    auto *rep = new VecSimQueryReply();
    last_mode_ = STANDARD_KNN;
    const size_t cap = std::min(k, indexSize());
    if (cap == 0) {
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

VecSimQueryReply *ShardedFlatIndex::rangeQuery(const void *blob, double radius, VecSimQueryParams *qp, VecSimQueryReply_Order order) {
    std::lock_guard<std::mutex> g(mu_);
    last_mode_ = RANGE_QUERY;
    std::vector<VecSimQueryReply *> parts(shards_.size(), nullptr);
    workers_->run([&](size_t sh) { parts[sh] = shards_[sh]->rangeQuery(blob, radius, qp, BY_SCORE); });
    auto *rep = new VecSimQueryReply();
    for (VecSimQueryReply *p : parts) {
        if (!p) continue;
        if (p->code != VecSim_QueryReply_OK) rep->code = p->code;
        rep->results.insert(rep->results.end(), p->results.begin(), p->results.end());
        delete p;
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

int ShardedFlatIndex::allScores(const void *processed_query, std::vector<std::pair<double, size_t>> &out) {
    std::lock_guard<std::mutex> g(mu_);
    std::vector<std::vector<std::pair<double, size_t>>> parts(shards_.size());
    std::vector<int> rcs(shards_.size(), 0);
    workers_->run([&](size_t sh) { rcs[sh] = shards_[sh]->allScores(processed_query, parts[sh]); });
    out.clear();
    for (size_t sh = 0; sh < shards_.size(); sh++) {
        if (rcs[sh] != 0) return -1;
        out.insert(out.end(), parts[sh].begin(), parts[sh].end());
    }
    return 0;
}

VecSimBatchIterator *ShardedFlatIndex::newBatchIterator(const void *blob, VecSimQueryParams *qp) {
    return new_flat_batch_iterator([this](const void *q, std::vector<std::pair<double, size_t>> &out) { return allScores(q, out); },
                                   indexLabelCount(), preprocessQuery(blob), qp ? qp->timeoutCtx : nullptr);
}

VecSimIndexBasicInfo ShardedFlatIndex::basicInfo() { return shards_[0]->basicInfo(); }

VecSimIndexStatsInfo ShardedFlatIndex::statsInfo() {
    VecSimIndexStatsInfo s{};
    s.memory = sizeof(*this) + ranges_.capacity() * sizeof(Range);
    for (auto &sh : shards_) s.memory += sh->statsInfo().memory;
    return s;
}

VecSimIndexDebugInfo ShardedFlatIndex::debugInfo() {
    VecSimIndexDebugInfo d{};
    d.commonInfo.basicInfo = basicInfo();
    d.commonInfo.indexSize = indexSize();
    d.commonInfo.indexLabelCount = indexLabelCount();
    d.commonInfo.memory = statsInfo().memory;
    d.commonInfo.lastMode = last_mode_;
    return d;
}

bool ShardedFlatIndex::preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) {
    // every shard scans its share at the same time: the batch pass costs what one shard's scan costs
    const bool res = prefer_adhoc_flat((indexSize() + shards_.size() - 1) / shards_.size(), stored_size_, subsetSize, k);
    last_mode_ = res ? (initial_check ? HYBRID_ADHOC_BF : HYBRID_BATCHES_TO_ADHOC_BF) : HYBRID_BATCHES;
    return res;
}

vsgpu_store *ShardedFlatIndex::deviceStore() { return shards_[0]->deviceStore(); }

void ShardedFlatIndex::lastStats(vsgpu_stats *out) {
    *out = vsgpu_stats{};
    for (auto &sh : shards_) {
        vsgpu_stats st{};
        sh->lastStats(&st);
        out->path = std::max(out->path, st.path);
        out->kernel_launches += st.kernel_launches;
        out->candidates += st.candidates;
        out->fallback_queries += st.fallback_queries;
        out->scan_ms = std::max(out->scan_ms, st.scan_ms);
    }
    out->total_ms = last_ms_;
}

} // namespace vsb
