/* oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY. Never linked into, loaded by, or called
 * from the product libraries (libvsgpu.so / libvecsim_b200.so).
 *
 * A plain-C façade over the UNMODIFIED reference sources compiled where they lie under
 * /root/reference (recipe: oracle/Makefile -> oracle/_ref/libvecsim_ref.so). It exists so that
 *   (1) the C restatement in oracle/vs_oracle.c can be pinned against the real thing, and
 *   (2) bench.py has the reference's own CPU path to time (cpu_baseline.kind = "reference").
 *
 * Reference entry points driven from here (paths relative to /root/reference/src/VecSim):
 *   index_factories/brute_force_factory.cpp:36-81  BruteForceFactory::NewIndex(const BFParams*)
 *   index_factories/hnsw_factory.cpp:37-77         HNSWFactory::NewIndex(const HNSWParams*)
 *   algorithms/brute_force/brute_force.h:242-326   topKQuery / rangeQuery
 *   algorithms/brute_force/bf_batch_iterator.h     getNextResults / isDepleted / reset
 *   algorithms/hnsw/hnsw.h:2037-2084               HNSWIndex::topKQuery
 *   spaces/IP_space.h, spaces/L2_space.h           *_GetDistFunc(dim, alignment, arch_opt)
 *   spaces/spaces.h:46-50                          GetNormalizeFunc<T>()
 * vec_sim.cpp itself is not compiled (it drags in SVS headers that need `fmt`), so the few
 * lines of glue it would provide (free-with-allocator, order handling) are restated here.
 */
#include "VecSim/vec_sim.h"
#include "VecSim/query_result_definitions.h"
#include "VecSim/batch_iterator.h"
#include "VecSim/index_factories/brute_force_factory.h"
#include "VecSim/index_factories/hnsw_factory.h"
#include "VecSim/algorithms/hnsw/hnsw.h"
#include "VecSim/spaces/spaces.h"
#include "VecSim/spaces/IP_space.h"
#include "VecSim/spaces/L2_space.h"
#include "VecSim/types/bfloat16.h"
#include "VecSim/types/float16.h"
#include "VecSim/utils/vec_utils.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

uint32_t vsref_feature_disable_mask = 0;

void vsref_set_feature_disable_mask(uint32_t mask) { vsref_feature_disable_mask = mask; }

/* Bitmask (shim numbering) of what the host CPU + OS actually offer. */
uint32_t vsref_host_features(void) {
    uint32_t saved = vsref_feature_disable_mask;
    vsref_feature_disable_mask = 0;
    auto f = cpu_features::GetX86Info().features;
    vsref_feature_disable_mask = saved;
    uint32_t m = 0;
    const int v[14] = {f.sse,     f.sse3,     f.sse4_1,   f.avx,        f.avx2,
                       f.fma3,    f.f16c,     f.avx512f,  f.avx512bw,   f.avx512vl,
                       f.avx512vnni, f.avx512vbmi2, f.avx512_bf16, f.avx512_fp16};
    for (int i = 0; i < 14; i++)
        if (v[i]) m |= 1u << i;
    return m;
}

void vsref_silence_logs(void) { VecSimIndexInterface::setLogCallbackFunction(nullptr); }

static int g_timeout_flag = 0;
static int timeout_cb(void *) { return g_timeout_flag; }
/* Install a timeout callback that returns `flag` (tests/unit/test_bruteforce.cpp:1489-1517). */
void vsref_set_timeout(int flag) {
    g_timeout_flag = flag;
    VecSimIndexInterface::setTimeoutCallbackFunction(timeout_cb); /* never nullptr: the reference calls it unchecked */
}

/* ---- distance kernels through the reference dispatcher ---------------------------------- */
/* type/metric use the VecSimType / VecSimMetric numbering. Returns 0 on success. */
int vsref_distance(int type, int metric, size_t dim, const void *a, const void *b, double *out) {
    unsigned char al = 0;
    try {
        switch (type) {
        case VecSimType_FLOAT32:
            *out = spaces::GetDistFunc<float, float>((VecSimMetric)metric, dim, &al)(a, b, dim);
            return 0;
        case VecSimType_FLOAT64:
            *out = spaces::GetDistFunc<double, double>((VecSimMetric)metric, dim, &al)(a, b, dim);
            return 0;
        case VecSimType_BFLOAT16:
            *out = spaces::GetDistFunc<vecsim_types::bfloat16, float>((VecSimMetric)metric, dim,
                                                                      &al)(a, b, dim);
            return 0;
        case VecSimType_FLOAT16:
            *out = spaces::GetDistFunc<vecsim_types::float16, float>((VecSimMetric)metric, dim,
                                                                     &al)(a, b, dim);
            return 0;
        case VecSimType_INT8:
            *out = spaces::GetDistFunc<int8_t, float>((VecSimMetric)metric, dim, &al)(a, b, dim);
            return 0;
        case VecSimType_UINT8:
            *out = spaces::GetDistFunc<uint8_t, float>((VecSimMetric)metric, dim, &al)(a, b, dim);
            return 0;
        }
    } catch (...) {
    }
    return -1;
}

/* Many pairs at once: out[i] = dist(a + i*stride_a, b + i*stride_b). */
int vsref_distance_many(int type, int metric, size_t dim, const void *a, size_t stride_a,
                        const void *b, size_t stride_b, size_t n, double *out) {
    for (size_t i = 0; i < n; i++) {
        int rc = vsref_distance(type, metric, dim, (const char *)a + i * stride_a,
                                (const char *)b + i * stride_b, out + i);
        if (rc) return rc;
    }
    return 0;
}

/* In-place normalisation exactly as VecSim_Normalize would do (vec_sim.cpp:230-254 picks
 * spaces::GetNormalizeFunc<T>()); int8/uint8 need dim+4 bytes of room (norm appended). */
int vsref_normalize(int type, size_t dim, void *blob) {
    switch (type) {
    case VecSimType_FLOAT32: spaces::GetNormalizeFunc<float>()(blob, dim); return 0;
    case VecSimType_FLOAT64: spaces::GetNormalizeFunc<double>()(blob, dim); return 0;
    case VecSimType_BFLOAT16: spaces::GetNormalizeFunc<vecsim_types::bfloat16>()(blob, dim); return 0;
    case VecSimType_FLOAT16: spaces::GetNormalizeFunc<vecsim_types::float16>()(blob, dim); return 0;
    case VecSimType_INT8: spaces::GetNormalizeFunc<int8_t>()(blob, dim); return 0;
    case VecSimType_UINT8: spaces::GetNormalizeFunc<uint8_t>()(blob, dim); return 0;
    }
    return -1;
}

/* ---- indexes ------------------------------------------------------------------------------ */
void *vsref_bf_new(int type, size_t dim, int metric, int multi, size_t block_size) {
    BFParams p{};
    p.type = (VecSimType)type;
    p.dim = dim;
    p.metric = (VecSimMetric)metric;
    p.multi = multi != 0;
    p.initialCapacity = 0;
    p.blockSize = block_size;
    try {
        return BruteForceFactory::NewIndex(&p);
    } catch (...) {
        return nullptr;
    }
}

void *vsref_hnsw_new(int type, size_t dim, int metric, int multi, size_t block_size, size_t M,
                     size_t ef_construction, size_t ef_runtime) {
    HNSWParams p{};
    p.type = (VecSimType)type;
    p.dim = dim;
    p.metric = (VecSimMetric)metric;
    p.multi = multi != 0;
    p.blockSize = block_size;
    p.M = M;
    p.efConstruction = ef_construction;
    p.efRuntime = ef_runtime;
    p.epsilon = HNSW_DEFAULT_EPSILON;
    try {
        return HNSWFactory::NewIndex(&p);
    } catch (...) {
        return nullptr;
    }
}

#ifdef BUILD_TESTS
/* Only in the BUILD_TESTS variant (make ref_bt): the reference's own index-file loader and writer
 * (index_factories/hnsw_factory.cpp:171-251, algorithms/hnsw/hnsw_serializer.cpp:39-52) and its integrity check. */
void *vsref_hnsw_load(const char *path) {
    try {
        return HNSWFactory::NewIndex(std::string(path));
    } catch (...) {
        return nullptr;
    }
}
int vsref_hnsw_save(void *h, const char *path) {
    try {
        auto *ser = dynamic_cast<HNSWSerializer *>((VecSimIndexInterface *)h);
        if (!ser) return -1;
        ser->saveIndex(std::string(path));
        return 0;
    } catch (...) {
        return -1;
    }
}
int vsref_hnsw_integrity(void *h, size_t *double_conn, size_t *unidir_conn) {
    auto *f32 = dynamic_cast<HNSWIndex<float, float> *>((VecSimIndexInterface *)h);
    if (!f32) return -1;
    HNSWIndexMetaData m = f32->checkIntegrity();
    if (double_conn) *double_conn = m.double_connections;
    if (unidir_conn) *unidir_conn = m.unidirectional_connections;
    return m.valid_state ? 1 : 0;
}
#endif

void vsref_index_free(void *h) {
    auto *idx = (VecSimIndexInterface *)h;
    if (!idx) return;
    /* vec_sim.cpp:371-375 — keep the allocator alive across the delete. */
    std::shared_ptr<VecSimAllocator> keep = idx->getAllocator();
    delete idx;
}

int vsref_add(void *h, const void *blob, size_t label) {
    return ((VecSimIndexInterface *)h)->addVector(blob, label);
}
/* Bulk ingest: labels[i] (or first_label+i when labels==NULL). Returns #new vectors. */
long vsref_add_many(void *h, const void *blobs, size_t stride, size_t n, const size_t *labels,
                    size_t first_label) {
    long added = 0;
    auto *idx = (VecSimIndexInterface *)h;
    for (size_t i = 0; i < n; i++)
        added += idx->addVector((const char *)blobs + i * stride, labels ? labels[i] : first_label + i);
    return added;
}
int vsref_delete(void *h, size_t label) { return ((VecSimIndexInterface *)h)->deleteVector(label); }
size_t vsref_size(void *h) { return ((VecSimIndexInterface *)h)->indexSize(); }
size_t vsref_label_count(void *h) { return ((VecSimIndexInterface *)h)->indexLabelCount(); }
double vsref_distance_from(void *h, size_t label, const void *blob) {
    return ((VecSimIndexInterface *)h)->getDistanceFrom_Unsafe(label, blob);
}

static size_t drain(VecSimQueryReply *rep, size_t cap, size_t *labels, double *scores, int *code) {
    size_t n = rep->results.size();
    if (code) *code = (int)rep->code;
    size_t m = n < cap ? n : cap;
    for (size_t i = 0; i < m; i++) {
        labels[i] = rep->results[i].id;
        scores[i] = rep->results[i].score;
    }
    delete rep;
    return n;
}

static void fill_qparams(VecSimQueryParams *qp, size_t ef_runtime) {
    memset(qp, 0, sizeof(*qp));
    qp->hnswRuntimeParams.efRuntime = ef_runtime;
}

/* order: BY_SCORE=0, BY_ID=1 (vec_sim.cpp:345-357). ef_runtime 0 = index default. */
size_t vsref_topk(void *h, const void *q, size_t k, int order, size_t ef_runtime, size_t *labels,
                  double *scores, int *code) {
    auto *idx = (VecSimIndexInterface *)h;
    VecSimQueryParams qp;
    fill_qparams(&qp, ef_runtime);
    VecSimQueryReply *rep = idx->topKQuery(q, k, &qp);
    if (order == BY_ID) sort_results_by_id(rep);
    return drain(rep, k, labels, scores, code);
}

/* Returns the total number of results (may exceed cap; only cap are written); -1 if the
 * reference threw (negative radius / bad order: vec_sim.cpp:362-367). */
long vsref_range(void *h, const void *q, double radius, int order, size_t cap, size_t *labels,
                 double *scores, int *code) {
    auto *idx = (VecSimIndexInterface *)h;
    if (order != BY_ID && order != BY_SCORE) return -1;
    if (radius < 0) return -1;
    try {
        VecSimQueryReply *rep = idx->rangeQuery(q, radius, nullptr, (VecSimQueryReply_Order)order);
        return (long)drain(rep, cap, labels, scores, code);
    } catch (...) {
        return -1;
    }
}

void *vsref_bi_new(void *h, const void *q) {
    return ((VecSimIndexInterface *)h)->newBatchIterator(q, nullptr);
}
size_t vsref_bi_next(void *it, size_t n, int order, size_t *labels, double *scores, int *code) {
    VecSimQueryReply *rep = ((VecSimBatchIterator *)it)->getNextResults(n, (VecSimQueryReply_Order)order);
    return drain(rep, n, labels, scores, code);
}
int vsref_bi_has_next(void *it) { return !((VecSimBatchIterator *)it)->isDepleted(); }
void vsref_bi_reset(void *it) { ((VecSimBatchIterator *)it)->reset(); }
void vsref_bi_free(void *it) { delete (VecSimBatchIterator *)it; }

/* The reference's own multi-query pattern (src/python_bindings/bindings.cpp:250-284): one
 * std::thread per core pulling query indices from an atomic counter, each a full topKQuery.
 * Writes nq*k labels/scores (padded with SIZE_MAX / NaN) and returns wall seconds. */
double vsref_topk_many(void *h, const void *queries, size_t stride, size_t nq, size_t k,
                       size_t ef_runtime, int n_threads, size_t *labels, double *scores) {
    auto *idx = (VecSimIndexInterface *)h;
    std::atomic<size_t> next{0};
    auto work = [&]() {
        VecSimQueryParams qp;
        fill_qparams(&qp, ef_runtime);
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= nq) break;
            VecSimQueryReply *rep = idx->topKQuery((const char *)queries + i * stride, k, &qp);
            size_t n = rep->results.size();
            for (size_t j = 0; j < k; j++) {
                if (labels) labels[i * k + j] = j < n ? rep->results[j].id : SIZE_MAX;
                if (scores) scores[i * k + j] = j < n ? rep->results[j].score : NAN;
            }
            delete rep;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (n_threads <= 1) {
        work();
    } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; t++) pool.emplace_back(work);
        for (auto &t : pool) t.join();
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

/* ---- HNSW graph export --------------------------------------------------- */
/* Only fp32 HNSW graphs are exported (config 5); the GPU search walks this exact graph. */
typedef HNSWIndex<float, float> HnswF32;

int vsref_hnsw_info(void *h, size_t *n, size_t *M, size_t *ef, long *entry, long *max_level) {
    auto *g = dynamic_cast<HnswF32 *>((VecSimIndexInterface *)h);
    if (!g) return -1;
    *n = g->indexSize();
    *M = g->getM();
    *ef = g->getEf();
    auto [ep, lvl] = g->safeGetEntryPointState();
    *entry = ep == INVALID_ID ? -1 : (long)ep;
    *max_level = ep == INVALID_ID ? -1 : (long)lvl;
    return 0;
}
/* levels[i] = top level of element i; labels[i]; flags[i] bit0 = deleted, bit1 = in-process. */
int vsref_hnsw_export_meta(void *h, uint32_t *levels, size_t *labels, uint8_t *flags) {
    auto *g = dynamic_cast<HnswF32 *>((VecSimIndexInterface *)h);
    if (!g) return -1;
    size_t n = g->indexSize();
    for (size_t i = 0; i < n; i++) {
        levels[i] = (uint32_t)g->getGraphDataByInternalId((idType)i)->toplevel;
        labels[i] = g->getExternalLabel((idType)i);
        flags[i] = (g->isMarkedDeleted((idType)i) ? 1 : 0) | (g->isInProcess((idType)i) ? 2 : 0);
    }
    return 0;
}
/* Links of every element at `level` into a dense [n x width] u32 table (UINT32_MAX padded),
 * counts[i] = number of links (0 for elements whose top level < level). */
int vsref_hnsw_export_level(void *h, size_t level, size_t width, uint32_t *links, uint32_t *counts) {
    auto *g = dynamic_cast<HnswF32 *>((VecSimIndexInterface *)h);
    if (!g) return -1;
    size_t n = g->indexSize();
    for (size_t i = 0; i < n; i++) {
        auto *el = g->getGraphDataByInternalId((idType)i);
        uint32_t *row = links + i * width;
        for (size_t j = 0; j < width; j++) row[j] = UINT32_MAX;
        if (el->toplevel < level) {
            counts[i] = 0;
            continue;
        }
        auto &ld = g->getElementLevelData(el, level);
        size_t c = ld.getNumLinks();
        if (c > width) return -2;
        counts[i] = (uint32_t)c;
        for (size_t j = 0; j < c; j++) row[j] = ld.getLinkAtPos(j);
    }
    return 0;
}
/* Stored (pre-processed) vector bytes of every element, packed at `stride`. */
int vsref_hnsw_export_vectors(void *h, size_t stride, size_t bytes, void *out) {
    auto *g = dynamic_cast<HnswF32 *>((VecSimIndexInterface *)h);
    if (!g) return -1;
    size_t n = g->indexSize();
    for (size_t i = 0; i < n; i++) memcpy((char *)out + i * stride, g->getDataByInternalId((idType)i), bytes);
    return 0;
}

} /* extern "C" */
