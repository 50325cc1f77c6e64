// extern "C" surface of libvecsim_b200.so: the reference's VecSimIndex_* / VecSimQueryReply_* /
// VecSimBatchIterator_* entry points (see include/vecsim_b200.h for the file:line each replaces)
// plus the batched / GPU additions. Thin: argument checks, ordering, object lifetime.
#include "vecsim_index.h"
#include "vecsim_hybrid.h"
#include "vecsim_numeric.h"
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <strings.h>

// ABI pins, measured from the reference headers with g++ 13 on x86-64 (SURVEY §8b "Struct ABI").
static_assert(sizeof(BFParams) == 40 && offsetof(BFParams, blockSize) == 32, "BFParams ABI");
static_assert(sizeof(HNSWParams) == 72 && offsetof(HNSWParams, epsilon) == 64, "HNSWParams ABI");
static_assert(sizeof(AlgoParams) == 120, "AlgoParams ABI");
static_assert(sizeof(TieredIndexParams) == 64 && offsetof(TieredIndexParams, submitCb) == 16 &&
                  offsetof(TieredIndexParams, primaryIndexParams) == 32 && offsetof(TieredIndexParams, specificParams) == 40,
              "TieredIndexParams ABI");
static_assert(sizeof(tieredInfoStruct) == 296 && offsetof(tieredInfoStruct, specificTieredBackendInfo) == 104 &&
                  offsetof(tieredInfoStruct, backendCommonInfo) == 136 && offsetof(tieredInfoStruct, frontendCommonInfo) == 200 &&
                  offsetof(tieredInfoStruct, bfInfo) == 264 && offsetof(tieredInfoStruct, management_layer_memory) == 272 &&
                  offsetof(tieredInfoStruct, backgroundIndexing) == 280 && offsetof(tieredInfoStruct, bufferLimit) == 288,
              "tieredInfoStruct ABI");
static_assert(sizeof(VecSimParams) == 136 && offsetof(VecSimParams, logCtx) == 128, "VecSimParams ABI");
static_assert(sizeof(VecSimQueryParams) == 56 && offsetof(VecSimQueryParams, batchSize) == 32 &&
                  offsetof(VecSimQueryParams, timeoutCtx) == 48, "VecSimQueryParams ABI");
static_assert(sizeof(VecSimIndexBasicInfo) == 32 && sizeof(VecSimIndexStatsInfo) == 32, "info ABI");
static_assert(sizeof(CommonInfo) == 64 && sizeof(VecSimIndexDebugInfo) == 360, "debug info ABI");
static_assert(sizeof(VecSimQueryResult) == 16 && sizeof(VecSimRawParam) == 32, "result ABI");

using namespace vsb;

static thread_local std::string g_api_err;

static void sort_reply(VecSimQueryReply *rep, VecSimQueryReply_Order order) {
    if (order == BY_ID)
        std::sort(rep->results.begin(), rep->results.end(),
                  [](const VecSimQueryResult &a, const VecSimQueryResult &b) { return a.id < b.id; });
    // BY_SCORE / BY_SCORE_THEN_ID: replies are produced in ascending (score, id) order already
}

extern "C" {

/* ---- replies ---- */
int64_t VecSimQueryResult_GetId(const VecSimQueryResult *item) { return item ? (int64_t)item->id : -1; }
double VecSimQueryResult_GetScore(const VecSimQueryResult *item) {
    return item ? item->score : std::numeric_limits<double>::quiet_NaN();
}
size_t VecSimQueryReply_Len(VecSimQueryReply *r) { return r->results.size(); }
VecSimQueryReply_Code VecSimQueryReply_GetCode(VecSimQueryReply *r) { return r->code; }
void VecSimQueryReply_Free(VecSimQueryReply *r) { delete r; }
VecSimQueryReply_Iterator *VecSimQueryReply_GetIterator(VecSimQueryReply *r) { return new VecSimQueryReply_Iterator{r, 0}; }
VecSimQueryResult *VecSimQueryReply_IteratorNext(VecSimQueryReply_Iterator *it) {
    if (it->pos >= it->reply->results.size()) return nullptr;
    return &it->reply->results[it->pos++];
}
bool VecSimQueryReply_IteratorHasNext(VecSimQueryReply_Iterator *it) { return it->pos < it->reply->results.size(); }
void VecSimQueryReply_IteratorReset(VecSimQueryReply_Iterator *it) { it->pos = 0; }
void VecSimQueryReply_IteratorFree(VecSimQueryReply_Iterator *it) { delete it; }

VecSimQueryReply *VecSimBatchIterator_Next(VecSimBatchIterator *it, size_t n, VecSimQueryReply_Order order) {
    return it->next(n, order);
}
bool VecSimBatchIterator_HasNext(VecSimBatchIterator *it) { return it->hasNext(); }
void VecSimBatchIterator_Free(VecSimBatchIterator *it) { delete it; }
void VecSimBatchIterator_Reset(VecSimBatchIterator *it) { it->reset(); }

/* ---- index ---- */
VecSimIndex *VecSimIndex_New(const VecSimParams *params) {
    // any failure -> NULL (index_factories/index_factory.cpp:40-43)
    try {
        if (!params) return nullptr;
        if (params->algo == VecSimAlgo_BF) {
            const BFParams &p = params->algoParams.bfParams;
            if (p.dim == 0 || p.type > VecSimType_UINT8 || p.metric > VecSimMetric_Cosine) return nullptr;
            if (p.multi) {
                auto *midx = new FlatMultiIndex(p, params->logCtx);
                if (!midx->ok()) {
                    g_api_err = std::string("device store: ") + vsgpu_last_error();
                    delete midx;
                    return nullptr;
                }
                return midx;
            }
            if (globals().devices.size() > 1 && p.type != VecSimType_FLOAT64) {
                // VecSimGPU_Configure named several devices: the rows are sharded over them (vecsim_flat_sharded.cpp)
                auto *sidx = new ShardedFlatIndex(p, params->logCtx, globals().devices);
                if (!sidx->ok()) {
                    g_api_err = std::string("sharded device stores: ") + vsgpu_last_error();
                    delete sidx;
                    return nullptr;
                }
                return sidx;
            }
            auto *idx = new FlatIndex(p, params->logCtx);
            if (!idx->ok()) {
                g_api_err = std::string("device store: ") + vsgpu_last_error();
                delete idx;
                return nullptr;
            }
            return idx;
        }
        if (params->algo == VecSimAlgo_HNSWLIB) {
            const HNSWParams &p = params->algoParams.hnswParams;
            if (p.dim == 0 || p.type > VecSimType_UINT8 || p.metric > VecSimMetric_Cosine) return nullptr;
            auto *idx = new HnswIndex(p, params->logCtx);
            if (!idx->ok()) {
                g_api_err = std::string("HNSW index: ") + vsgpu_last_error();
                delete idx;
                return nullptr;
            }
            return idx;
        }
        if (params->algo == VecSimAlgo_TIERED) {
            // index_factories/tiered_factory.cpp: only an HNSW backend (single value per label here)
            auto *idx = new TieredIndex(params->algoParams.tieredParams, params->logCtx);
            if (!idx->ok()) {
                const TieredIndexParams &tp = params->algoParams.tieredParams;
                const bool shape_ok = tp.primaryIndexParams && tp.primaryIndexParams->algo == VecSimAlgo_HNSWLIB;
                g_api_err = shape_ok ? std::string("tiered index: ") + vsgpu_last_error()
                                     : std::string("tiered index: the backend must be a VecSimAlgo_HNSWLIB index");
                delete idx;
                return nullptr;
            }
            return idx;
        }
        g_api_err = "VecSimAlgo_BF, VecSimAlgo_HNSWLIB and VecSimAlgo_TIERED are served by this library (not SVS)";
        return nullptr;
    } catch (...) {
        return nullptr;
    }
}
// What an empty index costs before the first vector (index_factories/*_factory.cpp EstimateInitialSize): the host object(s)
// plus what the constructor reserves on the device for `initialCapacity` rows (the reference ignores initialCapacity; here
// it reserves HBM rows up front, so it is part of the estimate).
size_t VecSimIndex_EstimateInitialSize(const VecSimParams *params) {
    if (!params) return 0;
    auto rows = [](VecSimType t, size_t dim, VecSimMetric m, size_t cap) {
        const size_t stride = (type_size(t) * dim + 15) / 16 * 16;
        const bool norm = m == VecSimMetric_Cosine && (t == VecSimType_INT8 || t == VecSimType_UINT8);
        return cap * (stride + sizeof(uint64_t) + (norm ? sizeof(float) : 0));
    };
    switch (params->algo) {
    case VecSimAlgo_BF: {
        const BFParams &p = params->algoParams.bfParams;
        return (p.multi ? sizeof(FlatMultiIndex) : sizeof(FlatIndex)) + rows(p.type, p.dim, p.metric, p.initialCapacity);
    }
    case VecSimAlgo_HNSWLIB: {
        const HNSWParams &p = params->algoParams.hnswParams;
        const size_t M = p.M ? p.M : 16;
        return sizeof(HnswIndex) + rows(p.type, p.dim, p.metric, p.initialCapacity) +
               p.initialCapacity * ((2 * M + 1) * sizeof(idType) + 3 * sizeof(uint32_t) + 1);
    }
    case VecSimAlgo_TIERED: {
        const TieredIndexParams &tp = params->algoParams.tieredParams;
        return sizeof(TieredIndex) + sizeof(FlatIndex) + (tp.primaryIndexParams ? VecSimIndex_EstimateInitialSize(tp.primaryIndexParams) : 0);
    }
    default: return 0;
    }
}
size_t VecSimIndex_EstimateElementSize(const VecSimParams *params) {
    if (params && params->algo == VecSimAlgo_HNSWLIB) {
        // index_factories/hnsw_factory.cpp:123-148: level-0 record + expected upper levels + vector + metadata
        const HNSWParams &p = params->algoParams.hnswParams;
        const size_t M = p.M ? p.M : 16;
        return stored_size(p.type, p.dim, p.metric) + (2 * M + 1) * sizeof(idType) + sizeof(size_t) + 2 * sizeof(idType) + 5;
    }
    if (!params || params->algo != VecSimAlgo_BF) return 0;
    const BFParams &p = params->algoParams.bfParams;
    return stored_size(p.type, p.dim, p.metric) + sizeof(size_t) + sizeof(idType);
}
void VecSimIndex_Free(VecSimIndex *index) { delete index; }
int VecSimIndex_AddVector(VecSimIndex *index, const void *blob, size_t label) { return index->addVector(blob, label); }
int VecSimIndex_DeleteVector(VecSimIndex *index, size_t label) { return index->deleteVector(label); }
double VecSimIndex_GetDistanceFrom_Unsafe(VecSimIndex *index, size_t label, const void *blob) {
    return index->getDistanceFrom(label, blob);
}
void VecSim_Normalize(void *blob, size_t dim, VecSimType type) { normalize_blob(blob, dim, type); }
size_t VecSimParams_GetQueryBlobSize(VecSimType type, size_t dim, VecSimMetric metric) { return stored_size(type, dim, metric); }
size_t VecSimIndex_IndexSize(VecSimIndex *index) { return index->indexSize(); }

static bool parse_positive(const VecSimRawParam &p, long long *out) {
    if (!p.value || p.valLen == 0) return false;
    char *end = nullptr;
    errno = 0;
    const long long v = strtoll(p.value, &end, 0);
    if (errno || end != p.value + p.valLen || v <= 0) return false;
    *out = v;
    return true;
}
static bool parse_positive_double(const VecSimRawParam &p, double *out) {
    if (!p.value || p.valLen == 0) return false;
    char *end = nullptr;
    errno = 0;
    const double v = strtod(p.value, &end);
    if (errno || end != p.value + p.valLen || !(v > 0) || std::isinf(v)) return false;
    *out = v;
    return true;
}

// vec_sim.cpp:270-343 restricted to the parameters a flat / HNSW index understands
VecSimResolveCode VecSimIndex_ResolveParams(VecSimIndex *index, VecSimRawParam *rparams, int paramNum,
                                            VecSimQueryParams *qparams, VecsimQueryType query_type) {
    if (!qparams || (!rparams && paramNum != 0)) return VecSimParamResolverErr_NullParam;
    const VecSimAlgo algo = index->basicInfo().algo;
    std::memset(qparams, 0, sizeof(*qparams));
    for (int i = 0; i < paramNum; i++) {
        const VecSimRawParam &p = rparams[i];
        long long iv;
        double dv;
        if (!strcasecmp(p.name, "EF_RUNTIME")) {
            if (algo != VecSimAlgo_HNSWLIB || query_type == QUERY_TYPE_RANGE) return VecSimParamResolverErr_UnknownParam;
            if (qparams->hnswRuntimeParams.efRuntime != 0) return VecSimParamResolverErr_AlreadySet;
            if (!parse_positive(p, &iv)) return VecSimParamResolverErr_BadValue;
            qparams->hnswRuntimeParams.efRuntime = (size_t)iv;
        } else if (!strcasecmp(p.name, "EPSILON")) {
            if (algo != VecSimAlgo_HNSWLIB) return VecSimParamResolverErr_UnknownParam;
            if (query_type != QUERY_TYPE_RANGE) return VecSimParamResolverErr_InvalidPolicy_NRange;
            if (qparams->hnswRuntimeParams.epsilon != 0) return VecSimParamResolverErr_AlreadySet;
            if (!parse_positive_double(p, &dv)) return VecSimParamResolverErr_BadValue;
            qparams->hnswRuntimeParams.epsilon = dv;
        } else if (!strcasecmp(p.name, "BATCH_SIZE")) {
            if (query_type != QUERY_TYPE_HYBRID) return VecSimParamResolverErr_InvalidPolicy_NHybrid;
            if (qparams->batchSize != 0) return VecSimParamResolverErr_AlreadySet;
            if (!parse_positive(p, &iv)) return VecSimParamResolverErr_BadValue;
            qparams->batchSize = (size_t)iv;
        } else if (!strcasecmp(p.name, "HYBRID_POLICY")) {
            if (query_type != QUERY_TYPE_HYBRID) return VecSimParamResolverErr_InvalidPolicy_NHybrid;
            if (qparams->searchMode != 0) return VecSimParamResolverErr_AlreadySet;
            if (p.value && !strcasecmp(p.value, "batches")) qparams->searchMode = HYBRID_BATCHES;
            else if (p.value && !strcasecmp(p.value, "adhoc_bf")) qparams->searchMode = HYBRID_ADHOC_BF;
            else return VecSimParamResolverErr_InvalidPolicy_NExits;
        } else {
            return VecSimParamResolverErr_UnknownParam;
        }
    }
    if (qparams->searchMode == HYBRID_ADHOC_BF && qparams->batchSize > 0)
        return VecSimParamResolverErr_InvalidPolicy_AdHoc_With_BatchSize;
    if (qparams->searchMode == HYBRID_ADHOC_BF && algo == VecSimAlgo_HNSWLIB && qparams->hnswRuntimeParams.efRuntime > 0)
        return VecSimParamResolverErr_InvalidPolicy_AdHoc_With_EfRuntime;
    if (qparams->searchMode != 0) index->setLastSearchMode(qparams->searchMode);
    return VecSimParamResolver_OK;
}

VecSimQueryReply *VecSimIndex_TopKQuery(VecSimIndex *index, const void *queryBlob, size_t k, VecSimQueryParams *queryParams,
                                        VecSimQueryReply_Order order) {
    VecSimQueryReply *rep = index->topKQuery(queryBlob, k, queryParams);
    sort_reply(rep, order);
    return rep;
}

VecSimQueryReply *VecSimIndex_RangeQuery(VecSimIndex *index, const void *queryBlob, double radius,
                                         VecSimQueryParams *queryParams, VecSimQueryReply_Order order) {
    // the reference throws through the C boundary here (vec_sim.cpp:362-367); kept for parity
    if (order != BY_ID && order != BY_SCORE && order != BY_SCORE_THEN_ID)
        throw std::runtime_error("Possible order values are only 'BY_ID' or 'BY_SCORE'");
    if (radius < 0) throw std::runtime_error("radius must be non-negative");
    return index->rangeQuery(queryBlob, radius, queryParams, order);
}

VecSimIndexDebugInfo VecSimIndex_DebugInfo(VecSimIndex *index) { return index->debugInfo(); }
VecSimIndexBasicInfo VecSimIndex_BasicInfo(VecSimIndex *index) { return index->basicInfo(); }

static const char *algo_str(VecSimAlgo a) {
    switch (a) {
    case VecSimAlgo_BF: return "FLAT";
    case VecSimAlgo_HNSWLIB: return "HNSW";
    case VecSimAlgo_TIERED: return "TIERED";
    default: return "SVS";
    }
}
static const char *type_str(VecSimType t) {
    static const char *n[] = {"FLOAT32", "FLOAT64", "BFLOAT16", "FLOAT16", "INT8", "UINT8", "INT32", "INT64"};
    return (unsigned)t < 8 ? n[t] : nullptr;
}
static const char *metric_str(VecSimMetric m) { return m == VecSimMetric_Cosine ? "COSINE" : (m == VecSimMetric_IP ? "IP" : "L2"); }
static const char *mode_str(VecSearchMode m) {
    static const char *n[] = {"EMPTY_MODE", "STANDARD_KNN", "HYBRID_ADHOC_BF", "HYBRID_BATCHES", "HYBRID_BATCHES_TO_ADHOC_BF", "RANGE_QUERY"};
    return (unsigned)m < 6 ? n[m] : nullptr;
}
// field names and order: brute_force.h:348-365, vec_sim_index.h:268-310 (common block), hnsw.h:2217-2272
VecSimDebugInfoIterator *VecSimIndex_DebugInfoIterator(VecSimIndex *index) {
    const VecSimIndexDebugInfo info = index->debugInfo();
    auto *it = new VecSimDebugInfoIterator();
    auto *tiered = dynamic_cast<TieredIndex *>(index);
    auto str = [&](const char *name, const char *v) {
        VecSim_InfoField f{};
        f.fieldName = name;
        f.fieldType = INFOFIELD_STRING;
        f.fieldValue.stringValue = v;
        it->fields.push_back(f);
    };
    auto u64 = [&](const char *name, uint64_t v) {
        VecSim_InfoField f{};
        f.fieldName = name;
        f.fieldType = INFOFIELD_UINT64;
        f.fieldValue.uintegerValue = v;
        it->fields.push_back(f);
    };
    const CommonInfo &c = info.commonInfo;
    if (tiered) {
        // vec_sim_tiered_index.h:393-441 + hnsw_tiered.h:1228-1241: the root names itself TIERED, then the common block,
        // the management fields, one nested iterator per tier and the HNSW-specific threshold
        str("ALGORITHM", "TIERED");
        str("TYPE", type_str(c.basicInfo.type));
        u64("DIMENSION", c.basicInfo.dim);
        str("METRIC", metric_str(c.basicInfo.metric));
        u64("IS_MULTI_VALUE", c.basicInfo.isMulti);
        u64("IS_DISK", c.basicInfo.isDisk);
        u64("INDEX_SIZE", c.indexSize);
        u64("INDEX_LABEL_COUNT", c.indexLabelCount);
        u64("MEMORY", c.memory);
        str("LAST_SEARCH_MODE", mode_str(c.lastMode));
        u64("MANAGEMENT_LAYER_MEMORY", info.tieredInfo.management_layer_memory);
        VecSim_InfoField bg{};
        bg.fieldName = "BACKGROUND_INDEXING";
        bg.fieldType = INFOFIELD_INT64;
        bg.fieldValue.integerValue = info.tieredInfo.backgroundIndexing;
        it->fields.push_back(bg);
        u64("TIERED_BUFFER_LIMIT", info.tieredInfo.bufferLimit);
        VecSim_InfoField fe{};
        fe.fieldName = "FRONTEND_INDEX";
        fe.fieldType = INFOFIELD_ITERATOR;
        fe.fieldValue.iteratorValue = VecSimIndex_DebugInfoIterator(tiered->frontend());
        it->fields.push_back(fe);
        VecSim_InfoField be{};
        be.fieldName = "BACKEND_INDEX";
        be.fieldType = INFOFIELD_ITERATOR;
        be.fieldValue.iteratorValue = VecSimIndex_DebugInfoIterator(tiered->backend());
        it->fields.push_back(be);
        u64("TIERED_HNSW_SWAP_JOBS_THRESHOLD", info.tieredInfo.specificTieredBackendInfo.hnswTieredInfo.pendingSwapJobsThreshold);
        return it;
    }
    str("ALGORITHM", algo_str(c.basicInfo.algo));
    str("TYPE", type_str(c.basicInfo.type));
    u64("DIMENSION", c.basicInfo.dim);
    str("METRIC", metric_str(c.basicInfo.metric));
    u64("IS_MULTI_VALUE", c.basicInfo.isMulti);
    u64("IS_DISK", c.basicInfo.isDisk);
    u64("INDEX_SIZE", c.indexSize);
    u64("INDEX_LABEL_COUNT", c.indexLabelCount);
    u64("MEMORY", c.memory);
    str("LAST_SEARCH_MODE", mode_str(c.lastMode));
    u64("BLOCK_SIZE", c.basicInfo.blockSize);
    if (c.basicInfo.algo == VecSimAlgo_HNSWLIB) {
        u64("M", info.hnswInfo.M);
        u64("EF_CONSTRUCTION", info.hnswInfo.efConstruction);
        u64("EF_RUNTIME", info.hnswInfo.efRuntime);
        u64("MAX_LEVEL", info.hnswInfo.max_level);
        u64("ENTRYPOINT", info.hnswInfo.entrypoint);
        VecSim_InfoField f{};
        f.fieldName = "EPSILON";
        f.fieldType = INFOFIELD_FLOAT64;
        f.fieldValue.floatingPointValue = info.hnswInfo.epsilon;
        it->fields.push_back(f);
        u64("NUMBER_OF_MARKED_DELETED", info.hnswInfo.numberOfMarkedDeletedNodes);
    }
    return it;
}
size_t VecSimDebugInfoIterator_NumberOfFields(VecSimDebugInfoIterator *it) { return it->fields.size(); }
bool VecSimDebugInfoIterator_HasNextField(VecSimDebugInfoIterator *it) { return it->pos < it->fields.size(); }
VecSim_InfoField *VecSimDebugInfoIterator_NextField(VecSimDebugInfoIterator *it) {
    return it->pos < it->fields.size() ? &it->fields[it->pos++] : nullptr;
}
void VecSimDebugInfoIterator_Free(VecSimDebugInfoIterator *it) {
    if (!it) return;
    for (auto &f : it->fields) // nested iterators belong to their parent (info_iterator.h:37-45)
        if (f.fieldType == INFOFIELD_ITERATOR) VecSimDebugInfoIterator_Free(f.fieldValue.iteratorValue);
    delete it;
}

int VecSimDebug_GetElementNeighborsInHNSWGraph(VecSimIndex *index, size_t label, int ***neighborsData) {
    *neighborsData = nullptr;
    if (index->basicInfo().algo != VecSimAlgo_HNSWLIB) return VecSimDebugCommandCode_BadIndex;
    return index->elementNeighbors(label, neighborsData);
}
void VecSimDebug_ReleaseElementNeighborsInHNSWGraph(int **neighborsData) {
    if (!neighborsData) return;
    for (size_t i = 0; neighborsData[i] != nullptr; i++) delete[] neighborsData[i];
    delete[] neighborsData;
}
VecSimIndexStatsInfo VecSimIndex_StatsInfo(VecSimIndex *index) { return index->statsInfo(); }
VecSimBatchIterator *VecSimBatchIterator_New(VecSimIndex *index, const void *queryBlob, VecSimQueryParams *queryParams) {
    return index->newBatchIterator(queryBlob, queryParams);
}
bool VecSimIndex_PreferAdHocSearch(VecSimIndex *index, size_t subsetSize, size_t k, bool initial_check) {
    return index->preferAdHocSearch(subsetSize, k, initial_check);
}

struct VecSimAdhocBfCtx {
    VecSimIndex *index;
    std::vector<uint8_t> query; // preprocessed once (vec_sim.h:240-274)
};
VecSimAdhocBfCtx *VecSimIndex_AdhocBfCtx_New(VecSimIndex *index, const void *queryBlob) {
    return new VecSimAdhocBfCtx{index, index->preprocessQuery(queryBlob)};
}
void VecSimIndex_AdhocBfCtx_Free(VecSimAdhocBfCtx *ctx) { delete ctx; }
double VecSimIndex_AdhocBfCtx_GetDistanceFrom(VecSimAdhocBfCtx *ctx, size_t label) {
    double d;
    ctx->index->exactDistances(ctx->query.data(), &label, &d, 1);
    return d;
}
void VecSimIndex_AdhocBfCtx_GetExactDistances(VecSimAdhocBfCtx *ctx, const size_t *labels, double *distances_out, size_t count) {
    ctx->index->exactDistances(ctx->query.data(), labels, distances_out, count);
}

// deleted backend nodes are tombstones here: there are no swap jobs to run
void VecSimTieredIndex_GC(VecSimIndex *) {}
void VecSimTieredIndex_AcquireSharedLocks(VecSimIndex *index) {
    if (auto *t = dynamic_cast<TieredIndex *>(index)) t->acquireSharedLocks();
}
void VecSimTieredIndex_ReleaseSharedLocks(VecSimIndex *index) {
    if (auto *t = dynamic_cast<TieredIndex *>(index)) t->releaseSharedLocks();
}
void VecSim_SetMemoryFunctions(VecSimMemoryFunctions f) {
    globals().mem = f;
    globals().mem_set = true;
}
void VecSim_SetTimeoutCallbackFunction(timeoutCallbackFunction cb) { globals().timeout_cb = cb; }
void VecSim_SetLogCallbackFunction(logCallbackFunction cb) { globals().log_cb = cb; }
void VecSim_SetTestLogContext(const char *, const char *) {}
void VecSim_SetWriteMode(VecSimWriteMode mode) { globals().write_mode = mode; }
void VecSim_UpdateThreadPoolSize(size_t new_size) { globals().write_mode = new_size == 0 ? VecSim_WriteInPlace : VecSim_WriteAsync; }
size_t VecSim_GetSharedMemory(void) { return 0; }

/* ---- additions ---- */
int VecSimIndex_TopKQueryBatchRaw(VecSimIndex *index, const void *queries, size_t nq, size_t k, VecSimQueryParams *qp,
                                  size_t *out_labels, double *out_scores) {
    const int rc = index->topKBatch(queries, nq, k, qp, out_labels, out_scores, nullptr);
    return rc == 0 ? 0 : (rc == 1 ? 1 : -1);
}

int VecSimIndex_TopKQueryBatch(VecSimIndex *index, const void *queries, size_t nq, size_t k, VecSimQueryParams *qp,
                               VecSimQueryReply_Order order, VecSimQueryReply **out) {
    if (nq == 0) return 0;
    const size_t kk = std::max<size_t>(std::min(k, index->indexSize()), 1);
    std::vector<size_t> labels(nq * kk);
    std::vector<double> scores(nq * kk);
    std::vector<uint32_t> counts(nq, 0);
    int rc = 0;
    if (k > 0 && index->indexSize() > 0) rc = index->topKBatch(queries, nq, kk, qp, labels.data(), scores.data(), counts.data());
    if (rc < 0) return -1;
    for (size_t q = 0; q < nq; q++) {
        auto *rep = new VecSimQueryReply();
        if (rc == 1) rep->code = VecSim_QueryReply_TimedOut;
        else {
            rep->results.resize(counts[q]);
            for (uint32_t j = 0; j < counts[q]; j++) rep->results[j] = {labels[q * kk + j], scores[q * kk + j]};
            sort_reply(rep, order);
        }
        out[q] = rep;
    }
    return 0;
}

long VecSimIndex_AddVectorBatch(VecSimIndex *index, const void *blobs, size_t n, const size_t *labels, size_t first_label) {
    return index->addVectorBatch(blobs, n, labels, first_label);
}

long VecSimGPU_AppendDeviceRows(VecSimIndex *index, const void *device_rows, size_t stride_bytes, size_t n, size_t first_label) {
    if (auto *sh = dynamic_cast<ShardedFlatIndex *>(index)) return sh->appendDeviceRows(device_rows, stride_bytes, n, first_label);
    auto *flat = dynamic_cast<FlatIndex *>(index);
    if (!flat) return -1;
    return flat->appendDeviceRows(device_rows, stride_bytes, n, first_label);
}

int VecSimGPU_SetDevice(int device) {
    if (device < 0 || device >= vsgpu_device_count()) return -1;
    globals().device = device;
    globals().devices.clear();
    return 0;
}
int VecSimGPU_Configure(const int *devices, size_t n) {
    if (!devices || n == 0) return -1;
    const int have = vsgpu_device_count();
    std::vector<int> d(devices, devices + n);
    for (size_t i = 0; i < n; i++)
        if (d[i] < 0 || d[i] >= have) return -1; // a device may be named more than once (several shards on one GPU)
    globals().device = d[0];
    globals().devices = n > 1 ? d : std::vector<int>();
    return 0;
}
size_t VecSimGPU_ShardCount(VecSimIndex *index) {
    auto *sh = dynamic_cast<ShardedFlatIndex *>(index);
    return sh ? sh->shardCount() : 1;
}
int VecSimGPU_GetDevice(void) { return globals().device; }
int VecSimGPU_DeviceCount(void) { return vsgpu_device_count(); }
void VecSimGPU_SetTopKMode(int mode) { globals().topk_mode = mode; }
void VecSimGPU_LastQueryStats(VecSimIndex *index, unsigned *path, unsigned *launches, uint64_t *candidates,
                              unsigned *fallbacks, float *scan_ms, float *total_ms) {
    vsgpu_stats st{};
    index->lastStats(&st);
    if (path) *path = st.path;
    if (launches) *launches = st.kernel_launches;
    if (candidates) *candidates = st.candidates;
    if (fallbacks) *fallbacks = st.fallback_queries;
    if (scan_ms) *scan_ms = st.scan_ms;
    if (total_ms) *total_ms = st.total_ms;
}
void *VecSimGPU_GetStore(VecSimIndex *index) { return index->deviceStore(); }
void *VecSimGPU_GetGraph(VecSimIndex *index) {
    auto *h = dynamic_cast<HnswIndex *>(index);
    return h ? (void *)h->deviceGraph() : nullptr;
}
int VecSimGPU_HNSWImportGraph(VecSimIndex *index, const void *blobs, int processed, size_t n, const size_t *labels,
                              const uint32_t *levels, const uint32_t *l0, const uint32_t *upper, size_t upper_records,
                              long entry, long max_level) {
    auto *h = dynamic_cast<HnswIndex *>(index);
    if (!h) return -1;
    return h->importGraph(blobs, processed, n, labels, levels, l0, upper, upper_records, entry, max_level);
}
int VecSimGPU_HNSWExportGraph(VecSimIndex *index, uint32_t *levels, uint32_t *l0, uint32_t *upper, size_t upper_cap_records,
                              size_t *upper_records, long *entry, long *max_level) {
    auto *h = dynamic_cast<HnswIndex *>(index);
    if (!h) return -1;
    vsgpu_hnsw *g = h->deviceGraph();
    if (!g) return -1;
    if (vsgpu_hnsw_export(g, levels, l0, upper, upper_cap_records, upper_records) != VSGPU_OK) return -1;
    vsgpu_hnsw_entry(g, entry, max_level);
    return 0;
}
int VecSimGPU_HNSWLastStats(VecSimIndex *index, unsigned long long *dist_evals, unsigned long long *hops, float *ms) {
    auto *h = dynamic_cast<HnswIndex *>(index);
    if (!h) return -1;
    return vsgpu_hnsw_last_stats(h->deviceGraph(), dist_evals, hops, ms);
}
VecSimIndex *VecSimGPU_HNSWLoadIndex(const char *path) {
    try {
        std::string err;
        VecSimIndexInterface *idx = path ? load_hnsw_file(path, err) : nullptr;
        if (!idx) g_api_err = err.empty() ? "VecSimGPU_HNSWLoadIndex: no path" : err;
        return idx;
    } catch (...) {
        g_api_err = "VecSimGPU_HNSWLoadIndex: out of memory or corrupted file";
        return nullptr;
    }
}
int VecSimGPU_HNSWSaveIndex(VecSimIndex *index, const char *path) {
    auto *h = dynamic_cast<HnswIndex *>(index);
    if (!h || !path) return -1;
    return h->saveFile(path);
}
const char *VecSimGPU_LastError(void) {
    if (!g_api_err.empty()) return g_api_err.c_str();
    return vsgpu_last_error();
}

/* host-logic hook for the CPU test-suite (no device needed): SURVEY App. A2 tie rule */
size_t vsb_test_resolve(const size_t *labels, const double *scores, const uint32_t *ids, size_t cnt, size_t k,
                        size_t *out_labels, double *out_scores) {
    std::vector<VecSimQueryResult> res;
    FlatIndex::resolve(labels, scores, ids, cnt, k, res);
    for (size_t i = 0; i < res.size(); i++) {
        out_labels[i] = res[i].id;
        out_scores[i] = res[i].score;
    }
    return res.size();
}

/* host-logic hook: the hybrid policy alone (algo 0 flat, 1 HNSW) */
int vsb_test_prefer_adhoc(int algo, size_t index_size, size_t row_bytes, size_t subset, size_t k) {
    return algo == 0 ? prefer_adhoc_flat(index_size, row_bytes, subset, k) : prefer_adhoc_hnsw(index_size, row_bytes, subset, k);
}

size_t vsb_test_tiered_merge(const size_t *a_ids, const double *a_scores, size_t na, const size_t *b_ids, const double *b_scores,
                             size_t nb, size_t limit, size_t *out_ids, double *out_scores, size_t *taken) {
    return tiered_merge_for_test(a_ids, a_scores, na, b_ids, b_scores, nb, limit, out_ids, out_scores, taken);
}

} // extern "C"
