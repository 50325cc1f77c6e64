// Tiered index: a flat buffer (front) that absorbs writes and an HNSW backend that background jobs drain it into —
// the caller RediSearch actually talks to. Host bookkeeping of the reference's
//   VecSimTieredIndex      /root/reference/src/VecSim/vec_sim_tiered_index.h:25-35 (job), :169-316 (top-K / range merge),
//                          :318-470 (stats / debug info)
//   TieredHNSWIndex        algorithms/hnsw/hnsw_tiered.h:576-636 (insert job), :742-939 (size / add / delete),
//                          :941-963 (getDistanceFrom), :975-1205 (batch iterator), :1208-1250 (info)
//   merge_results          utils/query_result_utils.h:14-126 (score-then-id order with a 1e-6 score epsilon)
// over the two device indexes of this library. Both tiers keep their rows in HBM and answer on the GPU; this file only
// orders the calls, merges the two replies and owns the insert jobs. There is no CPU distance code here.
//
// What is B200-shaped rather than mirrored: a job carries its vector, and the first job a worker runs moves EVERY vector
// pending at that moment into the backend as one device batch (one store append + one builder launch sequence instead of
// one per vector); the jobs of the vectors it took find themselves done when their turn comes. Deleting from the backend
// tombstones the node (as the reference's async mode does before its repair/swap jobs run); there are no repair or swap
// jobs, so VecSimTieredIndex_GC has nothing to do.
#include "vecsim_index.h"
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <shared_mutex>
#include <unordered_set>

struct AsyncJob {
    int jobType; // 0: insert into the backend (HNSW_INSERT_VECTOR_JOB)
    JobCallback Execute;
    VecSimIndex *index;
    bool isValid; // false once the vector was deleted / overwritten in the flat buffer, or already moved by another job
    size_t label;
    std::vector<uint8_t> blob;                // the caller's blob (both tiers preprocess it the same way)
    std::shared_ptr<std::atomic<bool>> alive; // false once the index is gone
    int attempts = 0;                         // failed ingestions of this vector so far
};

namespace vsb {

namespace {
constexpr double SCORE_EPS = 1e-6; // VECSIM_EPSILON, query_result_utils.h:14
inline int cmp_score_then_id(const VecSimQueryResult &a, const VecSimQueryResult &b) {
    // query_result_utils.h:18-23, including its (int) cast of the id difference
    return !(std::fabs(a.score - b.score) < SCORE_EPS) ? (a.score > b.score ? 1 : -1) : (int)(a.id - b.id);
}
inline bool less_score_id(const VecSimQueryResult &a, const VecSimQueryResult &b) {
    if (a.score < b.score) return true;
    if (b.score < a.score) return false;
    return a.id < b.id;
}

// merge_results<withSet = false> (query_result_utils.h:44-92): both inputs ascending by (score, id); a result present in
// both lists has the same score in both, meets itself during the merge and is emitted once.
// with_set = merge_results<withSet = true> (multi-value tiers: a label may sit in both lists with DIFFERENT scores; the
// better one comes first and the later one is dropped by the set of ids already emitted).
std::pair<size_t, size_t> merge_results(std::vector<VecSimQueryResult> &out, const std::vector<VecSimQueryResult> &first,
                                        const std::vector<VecSimQueryResult> &second, size_t limit, bool with_set = false) {
    out.reserve(std::min(limit, first.size() + second.size()));
    std::unordered_set<size_t> seen;
    auto append = [&](const VecSimQueryResult &r) {
        if (!with_set || seen.insert(r.id).second) {
            out.push_back(r);
            limit--;
        }
    };
    size_t i = 0, j = 0;
    while (limit && i < first.size() && j < second.size()) {
        const int c = cmp_score_then_id(first[i], second[j]);
        if (c > 0) append(second[j++]);
        else if (c < 0) append(first[i++]);
        else { // an exact duplicate: it did not appear before and will not appear again
            out.push_back(first[i]);
            i++;
            j++;
            limit--;
        }
    }
    if (limit != 0) {
        if (i == first.size())
            while (limit && j < second.size()) append(second[j++]);
        else
            while (limit && i < first.size()) append(first[i++]);
    }
    return {i, j};
}

// filter_results_by_id (query_result_utils.h:138-180): sort by id, keep one of each — the better score of a label that
// both tiers returned (multi-value); equal scores otherwise
void unique_by_id(std::vector<VecSimQueryResult> &r) {
    std::sort(r.begin(), r.end(), [](const VecSimQueryResult &a, const VecSimQueryResult &b) {
        if (a.id != b.id) return a.id < b.id;
        return a.score < b.score;
    });
    r.erase(std::unique(r.begin(), r.end(), [](const VecSimQueryResult &a, const VecSimQueryResult &b) { return a.id == b.id; }),
            r.end());
}
} // namespace

// host-logic hook for the CPU test-suite (tests/test_host_logic.py): the merge alone, no device needed
size_t tiered_merge_for_test(const size_t *a_ids, const double *a_scores, size_t na, const size_t *b_ids, const double *b_scores,
                             size_t nb, size_t limit, size_t *out_ids, double *out_scores, size_t *taken) {
    std::vector<VecSimQueryResult> a(na), b(nb), out;
    for (size_t i = 0; i < na; i++) a[i] = {a_ids[i], a_scores[i]};
    for (size_t i = 0; i < nb; i++) b[i] = {b_ids[i], b_scores[i]};
    const auto t = merge_results(out, a, b, limit);
    for (size_t i = 0; i < out.size(); i++) {
        out_ids[i] = out[i].id;
        out_scores[i] = out[i].score;
    }
    if (taken) {
        taken[0] = t.first;
        taken[1] = t.second;
    }
    return out.size();
}

TieredIndex::TieredIndex(const TieredIndexParams &tp, void *logCtx)
    : job_queue_(tp.jobQueue), job_queue_ctx_(tp.jobQueueCtx), submit_(tp.submitCb), flat_limit_(tp.flatBufferLimit),
      swap_threshold_(tp.specificParams.tieredHnswParams.swapJobThreshold), alive_(std::make_shared<std::atomic<bool>>(true)) {
    if (!tp.primaryIndexParams || tp.primaryIndexParams->algo != VecSimAlgo_HNSWLIB) return;
    const HNSWParams &hp = tp.primaryIndexParams->algoParams.hnswParams;
    multi_ = hp.multi;
    if (swap_threshold_ == 0) swap_threshold_ = 1024; // DEFAULT_PENDING_SWAP_JOBS_THRESHOLD (tiered_factory.cpp)
    auto *h = new HnswIndex(hp, logCtx);
    if (!h->ok()) {
        delete h;
        return;
    }
    back_.reset(h);
    // tiered_factory.cpp: the flat buffer shares type / dim / metric / blockSize with the backend
    BFParams bp{};
    bp.type = hp.type;
    bp.dim = hp.dim;
    bp.metric = hp.metric;
    bp.multi = multi_;
    bp.initialCapacity = 0;
    bp.blockSize = hp.blockSize;
    if (multi_) {
        auto *f = new FlatMultiIndex(bp, logCtx);
        if (!f->ok()) {
            delete f;
            back_.reset();
            return;
        }
        front_.reset(f);
    } else {
        auto *f = new FlatIndex(bp, logCtx);
        if (!f->ok()) {
            delete f;
            back_.reset();
            return;
        }
        front_.reset(f);
    }
    data_size_ = type_size(hp.type) * hp.dim;
}

TieredIndex::~TieredIndex() {
    alive_->store(false);
    for (AsyncJob *p : parked_) delete p; // never handed to the queue
    // jobs the queue still holds free themselves when they run; the ones it will never run are the caller's to drop
}

// ---- jobs ------------------------------------------------------------------------------------------------------------
void TieredIndex::executeJobWrapper(AsyncJob *job) {
    if (job->alive->load()) static_cast<TieredIndex *>(job->index)->executeInsertJob(job);
    delete job;
}

// hnsw_tiered.h:576-636, batched: everything pending goes to the backend in one device insertion
void TieredIndex::executeInsertJob(AsyncJob *job) {
    std::lock_guard<std::mutex> drain(drain_mu_);
    std::vector<AsyncJob *> batch;
    bool ok = true;
    size_t dropped = 0;
    {
        // the flat guard is held (shared) until the labels are registered in the backend: an overwrite / delete of one of
        // them either invalidated its job before this point or finds the label in the backend afterwards
        std::shared_lock<std::shared_mutex> flat(flat_guard_);
        if (!job->isValid) return;
        for (AsyncJob *p : pending_)
            if (p->isValid) batch.push_back(p);
        std::unique_lock<std::shared_mutex> main(main_guard_);
        for (AsyncJob *p : batch)
            if (back_->addVector(p->blob.data(), p->label) < 0) {
                ok = false;
                break;
            }
        flat.unlock();
        // store append + graph insertion on the device, under the exclusive main guard
        if (!ok || back_->sync() != 0) {
            ok = false;
            dropped = back_->abortPending(); // whatever did not reach the device stays in the flat buffer only
            if (globals().log_cb) {
                const std::string msg = "tiered index: the backend refused " + std::to_string(dropped) + " vector(s) (" +
                                        vsgpu_last_error() + "); they stay in the flat buffer";
                globals().log_cb(nullptr, "warning", msg.c_str());
            }
        }
    }
    // the backend takes vectors in order: the first (batch - dropped) of them are in, the rest are not
    const size_t ingested = ok ? batch.size() : (dropped >= batch.size() ? 0 : batch.size() - dropped);
    AsyncJob *retry = nullptr;
    {
        std::unique_lock<std::shared_mutex> flat(flat_guard_);
        for (size_t bi = 0; bi < batch.size(); bi++) {
            AsyncJob *p = batch[bi];
            if (!p->isValid) continue; // deleted / overwritten while it was being ingested: already out of the flat buffer
            if (bi < ingested) {
                // out of the flat buffer: the label's only vector (single value), or its oldest one (multi value: the vectors
                // of a label are ingested in the order they were added)
                if (multi_) static_cast<FlatMultiIndex *>(front_.get())->deleteFirst(p->label, 1);
                else front_->deleteVector(p->label);
                auto it = label_to_job_.find(p->label);
                if (it != label_to_job_.end()) {
                    auto &v = it->second;
                    v.erase(std::remove(v.begin(), v.end(), p), v.end());
                    if (v.empty()) label_to_job_.erase(it);
                }
                p->isValid = false;
                continue;
            }
            // not ingested: the vector stays buffered with a valid job. The job that is running dies when it returns, so its
            // vector moves to a fresh one — resubmitted twice, then parked (searchable in the buffer, never retried)
            p->attempts++;
            if (p != job) continue;
            auto *nj = new AsyncJob{0, executeJobWrapper, this, true, p->label, p->blob, alive_, p->attempts};
            auto &v = label_to_job_[p->label];
            std::replace(v.begin(), v.end(), p, nj);
            std::replace(pending_.begin(), pending_.end(), p, nj);
            p->isValid = false;
            if (nj->attempts < 3) retry = nj;
            else parked_.push_back(nj);
        }
        pending_.erase(std::remove_if(pending_.begin(), pending_.end(), [](AsyncJob *p) { return !p->isValid; }), pending_.end());
    }
    if (retry) submit_(job_queue_, job_queue_ctx_, &retry, &retry->Execute, 1);
}

void TieredIndex::invalidateJobLocked(size_t label) {
    auto it = label_to_job_.find(label);
    if (it == label_to_job_.end()) return;
    for (AsyncJob *j : it->second) { // one job per buffered vector of the label (several under a multi-value backend)
        j->isValid = false;
        // the job object stays with the queue (it frees itself when run); drop our references
        pending_.erase(std::remove(pending_.begin(), pending_.end(), j), pending_.end());
    }
    label_to_job_.erase(it);
}

// ---- writes ----------------------------------------------------------------------------------------------------------
int TieredIndex::addVector(const void *blob, size_t label) {
    int ret = 1;
    if (multi_) {
        // multi-value (hnsw_tiered.h:762-861 with HNSWIndex_Multi): a repeated label is another vector, never an overwrite
        if (globals().write_mode == VecSim_WriteInPlace || !submit_ || front_->indexSize() >= flat_limit_) {
            std::unique_lock<std::shared_mutex> main(main_guard_);
            if (back_->addVector(blob, label) < 0) return -1;
            ++direct_insertions_;
            return 1;
        }
        AsyncJob *job = nullptr;
        {
            std::unique_lock<std::shared_mutex> flat(flat_guard_);
            if (front_->addVector(blob, label) < 0) return -1;
            job = new AsyncJob{0, executeJobWrapper, this, true, label, {}, alive_};
            job->blob.assign((const uint8_t *)blob, (const uint8_t *)blob + data_size_);
            label_to_job_[label].push_back(job);
            pending_.push_back(job);
        }
        submit_(job_queue_, job_queue_ctx_, &job, &job->Execute, 1);
        return 1;
    }
    if (globals().write_mode == VecSim_WriteInPlace || !submit_) {
        ret -= deleteVector(label);
        std::unique_lock<std::shared_mutex> main(main_guard_);
        if (back_->addVector(blob, label) < 0) return -1;
        ++direct_insertions_;
        return std::max(ret, 0);
    }
    if (front_->indexSize() >= flat_limit_) {
        ret -= deleteVector(label);
        if (front_->indexSize() >= flat_limit_) {
            std::unique_lock<std::shared_mutex> main(main_guard_);
            if (back_->addVector(blob, label) < 0) return -1;
            ++direct_insertions_;
            return std::max(ret, 0);
        }
    }
    AsyncJob *job = nullptr;
    {
        std::unique_lock<std::shared_mutex> flat(flat_guard_);
        if (label_to_job_.count(label)) {
            invalidateJobLocked(label); // overwrite: the pending job of the old vector is void
            ret = 0;
        }
        if (front_->addVector(blob, label) < 0) return -1;
        job = new AsyncJob{0, executeJobWrapper, this, true, label, {}, alive_};
        job->blob.assign((const uint8_t *)blob, (const uint8_t *)blob + data_size_);
        label_to_job_[label] = {job};
        pending_.push_back(job);
    }
    // a worker may have ingested the previous vector of this label in the meantime: remove it from the backend before
    // the new job is submitted (hnsw_tiered.h:838-845)
    {
        std::unique_lock<std::shared_mutex> main(main_guard_);
        ret = std::max(ret - back_->deleteVector(label), 0);
    }
    submit_(job_queue_, job_queue_ctx_, &job, &job->Execute, 1);
    return ret;
}

long TieredIndex::addVectorBatch(const void *blobs, size_t n, const size_t *labels, size_t first_label) {
    long added = 0;
    for (size_t i = 0; i < n; i++) {
        const int rc = addVector((const uint8_t *)blobs + i * data_size_, labels ? labels[i] : first_label + i);
        if (rc < 0) return -1;
        added += rc;
    }
    return added;
}

int TieredIndex::deleteVector(size_t label) {
    int n = 0;
    {
        std::unique_lock<std::shared_mutex> flat(flat_guard_);
        if (label_to_job_.count(label)) {
            invalidateJobLocked(label);
            n += front_->deleteVector(label);
        }
    }
    // the same vector may be in the backend too if it was being ingested at that time (hnsw_tiered.h:897-913)
    std::unique_lock<std::shared_mutex> main(main_guard_);
    n += back_->deleteVector(label);
    return n;
}

// ---- reads -----------------------------------------------------------------------------------------------------------
size_t TieredIndex::indexSize() {
    std::shared_lock<std::shared_mutex> flat(flat_guard_);
    std::shared_lock<std::shared_mutex> main(main_guard_);
    return back_->indexSize() + front_->indexSize();
}

size_t TieredIndex::indexLabelCount() {
    std::shared_lock<std::shared_mutex> flat(flat_guard_);
    std::shared_lock<std::shared_mutex> main(main_guard_);
    size_t n = back_->indexLabelCount();
    for (const auto &kv : label_to_job_)
        if (!back_->hasLabel(kv.first)) n++;
    return n;
}

double TieredIndex::getDistanceFrom(size_t label, const void *blob) {
    const double d = front_->getDistanceFrom(label, blob);
    if (!multi_) {
        if (!std::isnan(d)) return d; // hnsw_tiered.h:941-963 (single value: the flat copy is authoritative)
        return back_->getDistanceFrom(label, blob);
    }
    const double b = back_->getDistanceFrom(label, blob); // multi value: the closest vector of the label in either tier
    if (std::isnan(d)) return b;
    if (std::isnan(b)) return d;
    return std::min(d, b);
}

void TieredIndex::exactDistances(const void *processed_query, const size_t *labels, double *out, size_t n) {
    front_->exactDistances(processed_query, labels, out, n);
    if (multi_) { // the closest vector of each label over both tiers
        std::vector<double> d(n);
        back_->exactDistances(processed_query, labels, d.data(), n);
        for (size_t i = 0; i < n; i++)
            if (std::isnan(out[i]) || d[i] < out[i]) out[i] = d[i];
        return;
    }
    std::vector<size_t> miss;
    std::vector<size_t> pos;
    for (size_t i = 0; i < n; i++)
        if (std::isnan(out[i])) {
            miss.push_back(labels[i]);
            pos.push_back(i);
        }
    if (miss.empty()) return;
    std::vector<double> d(miss.size());
    back_->exactDistances(processed_query, miss.data(), d.data(), miss.size());
    for (size_t j = 0; j < miss.size(); j++) out[pos[j]] = d[j];
}

// vec_sim_tiered_index.h:169-222
VecSimQueryReply *TieredIndex::topKQuery(const void *blob, size_t k, VecSimQueryParams *qp) {
    std::shared_lock<std::shared_mutex> flat(flat_guard_);
    if (front_->indexSize() == 0) {
        flat.unlock();
        std::shared_lock<std::shared_mutex> main(main_guard_);
        return back_->topKQuery(blob, k, qp);
    }
    VecSimQueryReply *flat_res = front_->topKQuery(blob, k, qp);
    flat.unlock();
    if (flat_res->code != VecSim_QueryReply_OK) return flat_res;
    VecSimQueryReply *main_res;
    {
        std::shared_lock<std::shared_mutex> main(main_guard_);
        main_res = back_->topKQuery(blob, k, qp);
    }
    if (main_res->code != VecSim_QueryReply_OK) {
        delete flat_res;
        return main_res;
    }
    std::sort(flat_res->results.begin(), flat_res->results.end(), less_score_id);
    std::sort(main_res->results.begin(), main_res->results.end(), less_score_id);
    auto *rep = new VecSimQueryReply();
    merge_results(rep->results, main_res->results, flat_res->results, k, multi_);
    delete flat_res;
    delete main_res;
    return rep;
}

int TieredIndex::topKBatch(const void *queries, size_t nq, size_t k, VecSimQueryParams *qp, size_t *labels, double *scores,
                           uint32_t *counts) {
    const double nan = std::numeric_limits<double>::quiet_NaN();
    std::vector<size_t> fl(nq * k), ml(nq * k);
    std::vector<double> fs(nq * k), ms(nq * k);
    std::vector<uint32_t> fc(nq, 0), mc(nq, 0);
    int rc = 0;
    {
        std::shared_lock<std::shared_mutex> flat(flat_guard_);
        if (front_->indexSize() && k) rc = front_->topKBatch(queries, nq, k, qp, fl.data(), fs.data(), fc.data());
    }
    if (rc == 0) {
        std::shared_lock<std::shared_mutex> main(main_guard_);
        if (back_->indexSize() && k) rc = back_->topKBatch(queries, nq, k, qp, ml.data(), ms.data(), mc.data());
    }
    if (rc != 0) {
        for (size_t i = 0; i < nq * k; i++) {
            if (labels) labels[i] = (size_t)-1;
            if (scores) scores[i] = nan;
        }
        if (counts) std::fill(counts, counts + nq, 0u);
        return rc;
    }
    std::vector<VecSimQueryResult> a, b, out;
    for (size_t q = 0; q < nq; q++) {
        a.clear();
        b.clear();
        out.clear();
        for (uint32_t j = 0; j < mc[q]; j++) a.push_back({ml[q * k + j], ms[q * k + j]});
        for (uint32_t j = 0; j < fc[q]; j++) b.push_back({fl[q * k + j], fs[q * k + j]});
        std::sort(a.begin(), a.end(), less_score_id);
        std::sort(b.begin(), b.end(), less_score_id);
        merge_results(out, a, b, k, multi_);
        for (size_t j = 0; j < k; j++) {
            if (labels) labels[q * k + j] = j < out.size() ? out[j].id : (size_t)-1;
            if (scores) scores[q * k + j] = j < out.size() ? out[j].score : nan;
        }
        if (counts) counts[q] = (uint32_t)out.size();
    }
    return 0;
}

// vec_sim_tiered_index.h:252-316
VecSimQueryReply *TieredIndex::rangeQuery(const void *blob, double radius, VecSimQueryParams *qp, VecSimQueryReply_Order order) {
    std::shared_lock<std::shared_mutex> flat(flat_guard_);
    if (front_->indexSize() == 0) {
        flat.unlock();
        std::shared_lock<std::shared_mutex> main(main_guard_);
        return back_->rangeQuery(blob, radius, qp, order);
    }
    VecSimQueryReply *flat_res = front_->rangeQuery(blob, radius, qp, BY_SCORE);
    flat.unlock();
    if (flat_res->code != VecSim_QueryReply_OK) return flat_res;
    VecSimQueryReply *main_res;
    {
        std::shared_lock<std::shared_mutex> main(main_guard_);
        main_res = back_->rangeQuery(blob, radius, qp, BY_SCORE);
    }
    auto *rep = new VecSimQueryReply();
    rep->code = main_res->code; // OK or timed out: the backend's code is the reply's
    if (order == BY_ID) {
        rep->results = std::move(main_res->results);
        rep->results.insert(rep->results.end(), flat_res->results.begin(), flat_res->results.end());
        unique_by_id(rep->results);
    } else {
        merge_results(rep->results, main_res->results, flat_res->results, (size_t)-1, multi_);
    }
    delete flat_res;
    delete main_res;
    return rep;
}

// TieredHNSW_BatchIterator (hnsw_tiered.h:975-1205), single-value variant
class TieredBatchIterator final : public VecSimBatchIterator {
  public:
    TieredBatchIterator(TieredIndex *idx, const void *blob, size_t blob_size, VecSimQueryParams *qp)
        : idx_(idx), blob_((const uint8_t *)blob, (const uint8_t *)blob + blob_size), has_qp_(qp != nullptr) {
        if (qp) qp_ = *qp;
        flat_it_ = idx_->front_->newBatchIterator(blob_.data(), qp);
    }
    ~TieredBatchIterator() override {
        delete flat_it_;
        releaseBackend();
    }
    VecSimQueryReply *next(size_t n, VecSimQueryReply_Order order) override {
        VecSimQueryReply_Code hnsw_code = VecSim_QueryReply_OK;
        if (state_ == UNINITIALIZED) {
            VecSimQueryReply *cur;
            {
                std::shared_lock<std::shared_mutex> flat(idx_->flat_guard_);
                cur = flat_it_->next(n, BY_SCORE_THEN_ID);
            }
            if (cur->code != VecSim_QueryReply_OK) return cur;
            flat_res_.swap(cur->results);
            delete cur;
            // the main guard is held (shared) from here until the backend iterator is depleted or freed
            idx_->main_guard_.lock_shared();
            hnsw_it_ = idx_->back_->newBatchIterator(blob_.data(), has_qp_ ? &qp_ : nullptr);
            state_ = ACTIVE;
            VecSimQueryReply *h = hnsw_it_->next(n, BY_SCORE_THEN_ID);
            hnsw_code = h->code;
            std::sort(h->results.begin(), h->results.end(), less_score_id);
            hnsw_res_.swap(h->results);
            delete h;
            if (!hnsw_it_->hasNext()) releaseBackend();
        } else {
            if (flat_res_.size() < n && flat_it_->hasNext()) {
                VecSimQueryReply *tail = flat_it_->next(n - flat_res_.size(), BY_SCORE_THEN_ID);
                flat_res_.insert(flat_res_.end(), tail->results.begin(), tail->results.end());
                delete tail;
                if (idx_->multi_) filterReturned(flat_res_); // multi value: a label the backend already gave out
            }
            while (hnsw_res_.size() < n && state_ == ACTIVE && hnsw_code == VecSim_QueryReply_OK) {
                VecSimQueryReply *tail = hnsw_it_->next(n - hnsw_res_.size(), BY_SCORE_THEN_ID);
                hnsw_code = tail->code;
                std::sort(tail->results.begin(), tail->results.end(), less_score_id);
                // a new batch may hold better results than what is left of the previous one
                std::vector<VecSimQueryResult> merged;
                merge_results(merged, hnsw_res_, tail->results, n, idx_->multi_);
                delete tail;
                hnsw_res_.swap(merged);
                filterReturned(hnsw_res_);
                if (!hnsw_it_->hasNext()) releaseBackend();
            }
        }
        auto *batch = new VecSimQueryReply();
        if (hnsw_code != VecSim_QueryReply_OK) {
            batch->code = hnsw_code;
            return batch;
        }
        const auto [from_hnsw, from_flat] = merge_results(batch->results, hnsw_res_, flat_res_, n, idx_->multi_);
        // what the flat buffer returned must not come back from the backend in a later batch
        for (size_t i = 0; i < from_flat; i++) returned_.insert(flat_res_[i].id);
        flat_res_.erase(flat_res_.begin(), flat_res_.begin() + (ptrdiff_t)from_flat);
        hnsw_res_.erase(hnsw_res_.begin(), hnsw_res_.begin() + (ptrdiff_t)from_hnsw);
        if (idx_->multi_) { // a label lives in both tiers with different vectors: whatever was handed out is done, in both
            for (const VecSimQueryResult &r : batch->results) returned_.insert(r.id);
            filterReturned(flat_res_);
            filterReturned(hnsw_res_);
        }
        if (order == BY_ID)
            std::sort(batch->results.begin(), batch->results.end(),
                      [](const VecSimQueryResult &a, const VecSimQueryResult &b) { return a.id < b.id; });
        return batch;
    }
    bool hasNext() override {
        const bool depleted = flat_res_.empty() && !flat_it_->hasNext() && hnsw_res_.empty() && state_ == DEPLETED;
        return !depleted;
    }
    void reset() override {
        releaseBackend();
        flat_it_->reset();
        state_ = UNINITIALIZED;
        flat_res_.clear();
        hnsw_res_.clear();
        returned_.clear();
    }

  private:
    enum State { UNINITIALIZED, ACTIVE, DEPLETED };
    void releaseBackend() {
        if (state_ == ACTIVE) {
            delete hnsw_it_;
            hnsw_it_ = nullptr;
            idx_->main_guard_.unlock_shared();
        }
        state_ = DEPLETED;
    }
    void filterReturned(std::vector<VecSimQueryResult> &r) {
        r.erase(std::remove_if(r.begin(), r.end(), [&](const VecSimQueryResult &x) { return returned_.count(x.id) != 0; }), r.end());
    }
    TieredIndex *idx_;
    std::vector<uint8_t> blob_;
    bool has_qp_;
    VecSimQueryParams qp_{};
    VecSimBatchIterator *flat_it_ = nullptr, *hnsw_it_ = nullptr;
    State state_ = UNINITIALIZED;
    std::vector<VecSimQueryResult> flat_res_, hnsw_res_;
    std::unordered_set<size_t> returned_;
};

VecSimBatchIterator *TieredIndex::newBatchIterator(const void *blob, VecSimQueryParams *qp) {
    return new TieredBatchIterator(this, blob, data_size_, qp);
}

// ---- info ------------------------------------------------------------------------------------------------------------
VecSimIndexBasicInfo TieredIndex::basicInfo() {
    VecSimIndexBasicInfo b = back_->basicInfo();
    b.isTiered = true;
    b.algo = VecSimAlgo_HNSWLIB; // hnsw_tiered.h:1244-1250
    return b;
}

VecSimIndexStatsInfo TieredIndex::statsInfo() {
    VecSimIndexStatsInfo s{};
    s.memory = sizeof(*this) + pending_.capacity() * sizeof(void *) + label_to_job_.size() * (48 + data_size_) +
               front_->statsInfo().memory + back_->statsInfo().memory;
    s.numberOfMarkedDeleted = back_->statsInfo().numberOfMarkedDeleted;
    s.directHNSWInsertions = direct_insertions_;
    s.flatBufferSize = front_->indexSize();
    return s;
}

VecSimIndexDebugInfo TieredIndex::debugInfo() {
    VecSimIndexDebugInfo info{};
    VecSimIndexDebugInfo f, b;
    {
        std::shared_lock<std::shared_mutex> flat(flat_guard_);
        std::shared_lock<std::shared_mutex> main(main_guard_);
        f = front_->debugInfo();
        b = back_->debugInfo();
    }
    info.commonInfo.indexLabelCount = indexLabelCount();
    info.commonInfo.indexSize = f.commonInfo.indexSize + b.commonInfo.indexSize;
    info.commonInfo.memory = statsInfo().memory;
    info.commonInfo.lastMode = b.commonInfo.lastMode;
    info.commonInfo.basicInfo = b.commonInfo.basicInfo;
    info.commonInfo.basicInfo.isTiered = true;
    info.tieredInfo.backendInfo.hnswInfo = b.hnswInfo;
    info.tieredInfo.specificTieredBackendInfo.hnswTieredInfo.pendingSwapJobsThreshold = swap_threshold_;
    info.tieredInfo.backendCommonInfo = b.commonInfo;
    info.tieredInfo.frontendCommonInfo = f.commonInfo;
    info.tieredInfo.bfInfo = f.bfInfo;
    info.tieredInfo.management_layer_memory = sizeof(*this) + pending_.capacity() * sizeof(void *) + label_to_job_.size() * (48 + data_size_);
    info.tieredInfo.backgroundIndexing = f.commonInfo.indexSize > 0 ? VecSimBool_TRUE : VecSimBool_FALSE;
    info.tieredInfo.bufferLimit = flat_limit_;
    return info;
}

bool TieredIndex::preferAdHocSearch(size_t subsetSize, size_t k, bool initial_check) {
    // decided by the bigger tier (vec_sim_tiered_index.h:149-154)
    return back_->indexSize() > front_->indexSize() ? back_->preferAdHocSearch(subsetSize, k, initial_check)
                                                    : front_->preferAdHocSearch(subsetSize, k, initial_check);
}

void TieredIndex::setLastSearchMode(VecSearchMode m) {
    front_->setLastSearchMode(m);
    back_->setLastSearchMode(m);
}
std::vector<uint8_t> TieredIndex::preprocessQuery(const void *blob) { return front_->preprocessQuery(blob); }
vsgpu_store *TieredIndex::deviceStore() { return back_->deviceStore(); }
void TieredIndex::lastStats(vsgpu_stats *out) { back_->lastStats(out); }
int TieredIndex::elementNeighbors(size_t label, int ***out) {
    std::shared_lock<std::shared_mutex> main(main_guard_);
    return back_->elementNeighbors(label, out);
}
void TieredIndex::acquireSharedLocks() {
    flat_guard_.lock_shared();
    main_guard_.lock_shared();
}
void TieredIndex::releaseSharedLocks() {
    main_guard_.unlock_shared();
    flat_guard_.unlock_shared();
}

} // namespace vsb
