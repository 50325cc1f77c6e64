/* include/vecsim_b200.h — the reference-facing C API of libvecsim_b200.so.
 *
 * Same symbols, argument meaning, ownership and error behaviour as the reference's public headers
 * for the flat (brute-force) path, so a consumer compiled against
 *   /root/reference/src/VecSim/vec_sim.h:28-331          (VecSimIndex_*, VecSim_Set*)
 *   /root/reference/src/VecSim/query_results.h:21-138    (VecSimQueryReply_*, VecSimBatchIterator_*)
 *   /root/reference/src/VecSim/vec_sim_common.h:60-475   (parameter / info structs, enums)
 * links against this library unchanged. Struct layouts are an ABI contract: every struct below is
 * byte-compatible with its namesake there (sizes and offsets are static_assert-ed in
 * vectorsimilarity_b200/csrc/host/vecsim_api.cpp against values measured from the reference).
 * Members this library never reads (SVS parameter blocks) are carried as opaque storage.
 *
 * New, non-breaking additions: VecSimIndex_TopKQueryBatch, VecSimGPU_*.
 */
#ifndef VECSIM_B200_H
#define VECSIM_B200_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums (vec_sim_common.h:60-117) ---- */
typedef enum {
    VecSimType_FLOAT32, VecSimType_FLOAT64, VecSimType_BFLOAT16, VecSimType_FLOAT16,
    VecSimType_INT8, VecSimType_UINT8, VecSimType_INT32, VecSimType_INT64
} VecSimType;
typedef enum { VecSimAlgo_BF, VecSimAlgo_HNSWLIB, VecSimAlgo_TIERED, VecSimAlgo_SVS } VecSimAlgo;
typedef enum { VecSimMetric_L2, VecSimMetric_IP, VecSimMetric_Cosine } VecSimMetric;
typedef enum { VecSimBool_TRUE = 1, VecSimBool_FALSE = 0, VecSimBool_UNSET = -1 } VecSimBool;
typedef size_t labelType;
typedef unsigned int idType;
#define VecSim_OK 0

typedef enum {
    VecSimParamResolver_OK = VecSim_OK, VecSimParamResolverErr_NullParam, VecSimParamResolverErr_AlreadySet,
    VecSimParamResolverErr_UnknownParam, VecSimParamResolverErr_BadValue,
    VecSimParamResolverErr_InvalidPolicy_NExits, VecSimParamResolverErr_InvalidPolicy_NHybrid,
    VecSimParamResolverErr_InvalidPolicy_NRange, VecSimParamResolverErr_InvalidPolicy_AdHoc_With_BatchSize,
    VecSimParamResolverErr_InvalidPolicy_AdHoc_With_EfRuntime
} VecSimResolveCode;
typedef enum { VecSim_WriteAsync, VecSim_WriteInPlace } VecSimWriteMode;
typedef enum { EMPTY_MODE, STANDARD_KNN, HYBRID_ADHOC_BF, HYBRID_BATCHES, HYBRID_BATCHES_TO_ADHOC_BF, RANGE_QUERY } VecSearchMode;
typedef enum { QUERY_TYPE_NONE, QUERY_TYPE_KNN, QUERY_TYPE_HYBRID, QUERY_TYPE_RANGE } VecsimQueryType;

/* ---- creation parameters (vec_sim_common.h:148-254) ---- */
typedef struct {
    VecSimType type; size_t dim; VecSimMetric metric; bool multi;
    size_t initialCapacity; /* deprecated in the reference; here: rows of HBM reserved up front */
    size_t blockSize;
} BFParams;
typedef struct {
    VecSimType type; size_t dim; VecSimMetric metric; bool multi;
    size_t initialCapacity; size_t blockSize;
    size_t M; size_t efConstruction; size_t efRuntime; double epsilon;
} HNSWParams;
typedef struct VecSimParams VecSimParams;
/* tiered index (vec_sim_common.h:126-141, 206-241): jobs are opaque to the caller, which only hands each job back to
 * its callback from one of its worker threads */
typedef struct AsyncJob AsyncJob;
typedef void (*JobCallback)(AsyncJob *);
typedef int (*SubmitCB)(void *job_queue, void *index_ctx, AsyncJob **jobs, JobCallback *CBs, size_t jobs_len);
typedef struct { size_t swapJobThreshold; } TieredHNSWParams;
typedef struct {
    void *jobQueue; void *jobQueueCtx; SubmitCB submitCb;
    size_t flatBufferLimit;            /* flat buffer full -> in-place insertion into the backend */
    VecSimParams *primaryIndexParams;  /* the backend index (HNSW) */
    union { TieredHNSWParams tieredHnswParams; uint64_t _opaque[3]; /* SVS / disk variants */ } specificParams;
} TieredIndexParams;
typedef union {
    HNSWParams hnswParams;
    BFParams bfParams;
    TieredIndexParams tieredParams;
    uint64_t _opaque[15]; /* SVSParams storage (120 bytes): never read here */
} AlgoParams;
struct VecSimParams {
    VecSimAlgo algo;
    AlgoParams algoParams;
    void *logCtx;
};

typedef struct { const char *name; size_t nameLen; const char *value; size_t valLen; } VecSimRawParam;

/* ---- per-query parameters (vec_sim_common.h:280-338) ---- */
typedef struct { size_t efRuntime; double epsilon; } HNSWRuntimeParams;
typedef struct {
    union {
        HNSWRuntimeParams hnswRuntimeParams;
        uint64_t _opaque[4]; /* HNSWDisk / SVS runtime params */
    };
    size_t batchSize;
    VecSearchMode searchMode;
    void *timeoutCtx;
} VecSimQueryParams;

/* ---- info structs (vec_sim_common.h:343-447) ---- */
typedef struct {
    VecSimAlgo algo; VecSimMetric metric; VecSimType type;
    bool isMulti; bool isTiered; bool isDisk;
    size_t blockSize; size_t dim;
} VecSimIndexBasicInfo;
typedef struct { size_t memory; size_t numberOfMarkedDeleted; size_t directHNSWInsertions; size_t flatBufferSize; } VecSimIndexStatsInfo;
typedef struct {
    VecSimIndexBasicInfo basicInfo; size_t indexSize; size_t indexLabelCount; uint64_t memory; VecSearchMode lastMode;
} CommonInfo;
typedef struct {
    size_t M, efConstruction, efRuntime; double epsilon; size_t max_level, entrypoint, visitedNodesPoolSize, numberOfMarkedDeletedNodes;
} hnswInfoStruct;
typedef struct { char dummy; } bfInfoStruct;
typedef struct { size_t pendingSwapJobsThreshold; } HnswTieredInfo;
typedef struct {
    union { hnswInfoStruct hnswInfo; uint64_t _opaque[13]; /* svsInfoStruct */ } backendInfo;
    union { HnswTieredInfo hnswTieredInfo; uint64_t _opaque[4]; /* SvsTieredInfo */ } specificTieredBackendInfo;
    CommonInfo backendCommonInfo; CommonInfo frontendCommonInfo; bfInfoStruct bfInfo;
    uint64_t management_layer_memory; VecSimBool backgroundIndexing; size_t bufferLimit;
} tieredInfoStruct;
typedef struct {
    CommonInfo commonInfo;
    union { bfInfoStruct bfInfo; hnswInfoStruct hnswInfo; tieredInfoStruct tieredInfo; uint64_t _opaque[37]; /* svs */ };
} VecSimIndexDebugInfo;

/* ---- callbacks (vec_sim_common.h:452-490) ---- */
typedef struct {
    void *(*allocFunction)(size_t n);
    void *(*callocFunction)(size_t nelem, size_t elemsz);
    void *(*reallocFunction)(void *p, size_t n);
    void (*freeFunction)(void *p);
} VecSimMemoryFunctions;
typedef int (*timeoutCallbackFunction)(void *ctx);
typedef void (*logCallbackFunction)(void *ctx, const char *level, const char *message);

/* ---- debug info iterator (info_iterator.h:21-81) and debug commands (vec_sim_debug.h, vec_sim_common.h:119-124) ---- */
typedef struct VecSimDebugInfoIterator VecSimDebugInfoIterator;
typedef enum { INFOFIELD_STRING, INFOFIELD_INT64, INFOFIELD_UINT64, INFOFIELD_FLOAT64, INFOFIELD_ITERATOR } VecSim_InfoFieldType;
typedef union {
    double floatingPointValue; int64_t integerValue; uint64_t uintegerValue; const char *stringValue;
    VecSimDebugInfoIterator *iteratorValue;
} FieldValue;
typedef struct { const char *fieldName; VecSim_InfoFieldType fieldType; FieldValue fieldValue; } VecSim_InfoField;
typedef enum {
    VecSimDebugCommandCode_OK = 0, VecSimDebugCommandCode_BadIndex, VecSimDebugCommandCode_LabelNotExists,
    VecSimDebugCommandCode_MultiNotSupported
} VecSimDebugCommandCode;

/* ---- replies (query_results.h:21-138) ---- */
typedef enum { BY_SCORE, BY_ID, BY_SCORE_THEN_ID } VecSimQueryReply_Order;
typedef enum { VecSim_QueryReply_OK = VecSim_OK, VecSim_QueryReply_TimedOut } VecSimQueryReply_Code;
typedef struct VecSimQueryResult VecSimQueryResult;
typedef struct VecSimQueryReply VecSimQueryReply;
typedef struct VecSimQueryReply_Iterator VecSimQueryReply_Iterator;
typedef struct VecSimBatchIterator VecSimBatchIterator;
typedef struct VecSimIndexInterface VecSimIndex;
typedef struct VecSimAdhocBfCtx VecSimAdhocBfCtx;

int64_t VecSimQueryResult_GetId(const VecSimQueryResult *item);     /* NULL -> INVALID_ID (-1) */
double VecSimQueryResult_GetScore(const VecSimQueryResult *item);   /* NULL -> NaN */
size_t VecSimQueryReply_Len(VecSimQueryReply *results);
VecSimQueryReply_Code VecSimQueryReply_GetCode(VecSimQueryReply *results);
void VecSimQueryReply_Free(VecSimQueryReply *results);
VecSimQueryReply_Iterator *VecSimQueryReply_GetIterator(VecSimQueryReply *results);
VecSimQueryResult *VecSimQueryReply_IteratorNext(VecSimQueryReply_Iterator *iterator);
bool VecSimQueryReply_IteratorHasNext(VecSimQueryReply_Iterator *iterator);
void VecSimQueryReply_IteratorReset(VecSimQueryReply_Iterator *iterator);
void VecSimQueryReply_IteratorFree(VecSimQueryReply_Iterator *iterator);

VecSimQueryReply *VecSimBatchIterator_Next(VecSimBatchIterator *iterator, size_t n_results, VecSimQueryReply_Order order);
bool VecSimBatchIterator_HasNext(VecSimBatchIterator *iterator);
void VecSimBatchIterator_Free(VecSimBatchIterator *iterator);
void VecSimBatchIterator_Reset(VecSimBatchIterator *iterator);

/* ---- index API (vec_sim.h:28-331) ---- */
VecSimIndex *VecSimIndex_New(const VecSimParams *params);           /* NULL on any failure */
size_t VecSimIndex_EstimateInitialSize(const VecSimParams *params);
size_t VecSimIndex_EstimateElementSize(const VecSimParams *params);
void VecSimIndex_Free(VecSimIndex *index);
int VecSimIndex_AddVector(VecSimIndex *index, const void *blob, size_t label);   /* #new vectors (0 = overwrite) */
int VecSimIndex_DeleteVector(VecSimIndex *index, size_t label);                  /* #deleted */
double VecSimIndex_GetDistanceFrom_Unsafe(VecSimIndex *index, size_t label, const void *blob);
void VecSim_Normalize(void *blob, size_t dim, VecSimType type);
size_t VecSimParams_GetQueryBlobSize(VecSimType type, size_t dim, VecSimMetric metric);
size_t VecSimIndex_IndexSize(VecSimIndex *index);
VecSimResolveCode VecSimIndex_ResolveParams(VecSimIndex *index, VecSimRawParam *rparams, int paramNum,
                                            VecSimQueryParams *qparams, VecsimQueryType query_type);
VecSimQueryReply *VecSimIndex_TopKQuery(VecSimIndex *index, const void *queryBlob, size_t k,
                                        VecSimQueryParams *queryParams, VecSimQueryReply_Order);
VecSimQueryReply *VecSimIndex_RangeQuery(VecSimIndex *index, const void *queryBlob, double radius,
                                         VecSimQueryParams *queryParams, VecSimQueryReply_Order);
VecSimIndexDebugInfo VecSimIndex_DebugInfo(VecSimIndex *index);
/* vec_sim.h:196-202, info_iterator.h:57-81: field names / order as the reference's debugInfoIterator()
 * (brute_force.h:348-365, hnsw.h:2217-2272). The caller frees the iterator. */
VecSimDebugInfoIterator *VecSimIndex_DebugInfoIterator(VecSimIndex *index);
size_t VecSimDebugInfoIterator_NumberOfFields(VecSimDebugInfoIterator *infoIterator);
bool VecSimDebugInfoIterator_HasNextField(VecSimDebugInfoIterator *infoIterator);
VecSim_InfoField *VecSimDebugInfoIterator_NextField(VecSimDebugInfoIterator *infoIterator);
void VecSimDebugInfoIterator_Free(VecSimDebugInfoIterator *infoIterator);
/* vec_sim_debug.h: per level, [count, neighbour labels...]; the array ends with a NULL row (hnsw.h:2414-2441) */
int VecSimDebug_GetElementNeighborsInHNSWGraph(VecSimIndex *index, size_t label, int ***neighborsData);
void VecSimDebug_ReleaseElementNeighborsInHNSWGraph(int **neighborsData);
VecSimIndexBasicInfo VecSimIndex_BasicInfo(VecSimIndex *index);
VecSimIndexStatsInfo VecSimIndex_StatsInfo(VecSimIndex *index);
VecSimBatchIterator *VecSimBatchIterator_New(VecSimIndex *index, const void *queryBlob, VecSimQueryParams *queryParams);
bool VecSimIndex_PreferAdHocSearch(VecSimIndex *index, size_t subsetSize, size_t k, bool initial_check);
VecSimAdhocBfCtx *VecSimIndex_AdhocBfCtx_New(VecSimIndex *index, const void *queryBlob);
void VecSimIndex_AdhocBfCtx_Free(VecSimAdhocBfCtx *ctx);
double VecSimIndex_AdhocBfCtx_GetDistanceFrom(VecSimAdhocBfCtx *ctx, size_t label);
void VecSimIndex_AdhocBfCtx_GetExactDistances(VecSimAdhocBfCtx *ctx, const size_t *labels, double *distances_out, size_t count);
void VecSimTieredIndex_GC(VecSimIndex *index);
void VecSimTieredIndex_AcquireSharedLocks(VecSimIndex *index);
void VecSimTieredIndex_ReleaseSharedLocks(VecSimIndex *index);
void VecSim_SetMemoryFunctions(VecSimMemoryFunctions memoryfunctions);
void VecSim_SetTimeoutCallbackFunction(timeoutCallbackFunction callback);
void VecSim_SetLogCallbackFunction(logCallbackFunction callback);
void VecSim_SetTestLogContext(const char *test_name, const char *test_type);
void VecSim_SetWriteMode(VecSimWriteMode mode);
void VecSim_UpdateThreadPoolSize(size_t new_size);
size_t VecSim_GetSharedMemory(void);

/* ---- additions ---- */
/* nq queries (`queries` = nq blobs of dim*sizeof(type) bytes, back to back) in one scan of the
 * store. out[i] receives what VecSimIndex_TopKQuery(index, query_i, k, queryParams, order) would
 * return; the caller frees each with VecSimQueryReply_Free. Returns 0, or -1 on a device error
 * (out[] untouched). */
int VecSimIndex_TopKQueryBatch(VecSimIndex *index, const void *queries, size_t nq, size_t k,
                               VecSimQueryParams *queryParams, VecSimQueryReply_Order order, VecSimQueryReply **out);
/* Same, results into caller arrays [nq][k] (labels SIZE_MAX / scores NaN padded): no per-reply
 * allocations; this is what benchmarks and the sharded front-end call. */
int VecSimIndex_TopKQueryBatchRaw(VecSimIndex *index, const void *queries, size_t nq, size_t k,
                                  VecSimQueryParams *queryParams, size_t *out_labels, double *out_scores);
/* Bulk ingest: n blobs back to back, labels[i] (NULL: first_label + i). Returns #new vectors or -1. */
long VecSimIndex_AddVectorBatch(VecSimIndex *index, const void *blobs, size_t n, const size_t *labels, size_t first_label);

/* Bulk ingest of rows that already sit in DEVICE memory of the index's GPU, in processed form
 * (normalised for Cosine; int8/uint8 Cosine: dim bytes per row, the norm is computed on the device).
 * Row i gets label first_label + i. Returns #rows added or -1. */
long VecSimGPU_AppendDeviceRows(VecSimIndex *index, const void *device_rows, size_t stride_bytes, size_t n, size_t first_label);

/* Device placement for indexes created afterwards (process-wide; default device 0). */
int VecSimGPU_SetDevice(int device);
int VecSimGPU_GetDevice(void);
int VecSimGPU_DeviceCount(void);
/* Several devices (SURVEY.md §5, §8b "Additions"): flat (single-value, non-fp64) indexes created afterwards keep their
 * rows sharded over `devices` — one process, one host thread and one stream per device, per-shard top-k lists gathered
 * over NVLink peer copies and merged on devices[0] — behind the unchanged VecSimIndex_* calls. n == 1 is
 * VecSimGPU_SetDevice. Labels are routed by (label mod n); rows bulk-ingested with VecSimGPU_AppendDeviceRows stay on the
 * device they were on (which must be one of `devices`). A device may be listed more than once. Returns 0, or -1 for an
 * unknown device. */
int VecSimGPU_Configure(const int *devices, size_t n);
size_t VecSimGPU_ShardCount(VecSimIndex *index); /* 1 for an index that lives on one device */
/* 0 auto, 1 exact scan only, 2 tensor path only (see include/vsgpu.h flags). */
void VecSimGPU_SetTopKMode(int mode);
/* Counters of the last query on this index: path (0 exact, 1 tensor), kernel launches, candidates,
 * fallback queries, scan ms, total ms. Any pointer may be NULL. */
void VecSimGPU_LastQueryStats(VecSimIndex *index, unsigned *path, unsigned *launches, uint64_t *candidates,
                              unsigned *fallbacks, float *scan_ms, float *total_ms);
/* The device store behind a flat index (vsgpu_store*, include/vsgpu.h) for callers that keep queries
 * and results on the device (sharded multi-GPU front-end). Flushes pending appends first. */
void *VecSimGPU_GetStore(VecSimIndex *index);
/* HNSW indexes: the device graph (vsgpu_hnsw*), bulk load of a graph built elsewhere over rows in
 * insertion order (`processed` != 0: blobs are stored rows, e.g. already normalised; layout of
 * levels / l0 / upper as vsgpu_hnsw_import), read-back, and counters of the last traversal. */
void *VecSimGPU_GetGraph(VecSimIndex *index);
int VecSimGPU_HNSWImportGraph(VecSimIndex *index, const void *blobs, int processed, size_t n, const size_t *labels,
                              const uint32_t *levels, const uint32_t *l0, const uint32_t *upper, size_t upper_records,
                              long entry, long max_level);
int VecSimGPU_HNSWExportGraph(VecSimIndex *index, uint32_t *levels, uint32_t *l0, uint32_t *upper, size_t upper_cap_records,
                              size_t *upper_records, long *entry, long *max_level);
/* HNSW index files in the reference's serialized format: load = HNSWFactory::NewIndex(location)
 * (index_factories/hnsw_factory.cpp:171-251; encoding V3 and V4, single value per label; NULL + VecSimGPU_LastError on a
 * bad file), save = HNSWIndex::saveIndex (algorithms/hnsw/hnsw_serializer_impl.h:247-330; writes encoding V4, which the
 * reference loads back). Rows land in the device store, links in the device graph. */
VecSimIndex *VecSimGPU_HNSWLoadIndex(const char *path);
int VecSimGPU_HNSWSaveIndex(VecSimIndex *index, const char *path);
int VecSimGPU_HNSWLastStats(VecSimIndex *index, unsigned long long *dist_evals, unsigned long long *hops, float *ms);
const char *VecSimGPU_LastError(void);

#ifdef __cplusplus
}
#endif
#endif
