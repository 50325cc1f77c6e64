"""ctypes view of libvecsim_b200.so — the C API a RediSearch-style consumer binds (include/vecsim_b200.h).

The Python classes mirror the reference's own Python binding (src/python_bindings/bindings.cpp:
BFParams, BFIndex.add_vector / delete_vector / knn_query / range_query / index_size /
create_batch_iterator, BatchIterator.has_next / get_next_results / reset) so the parity tests read
like the reference's tests/flow suite. Every call goes through the C-ABI; there is no Python or CPU
fallback: loading fails loudly when the CUDA libraries are missing.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# enums (include/vecsim_b200.h)
VecSimType_FLOAT32, VecSimType_FLOAT64, VecSimType_BFLOAT16, VecSimType_FLOAT16, VecSimType_INT8, \
    VecSimType_UINT8 = range(6)
VecSimAlgo_BF, VecSimAlgo_HNSWLIB, VecSimAlgo_TIERED, VecSimAlgo_SVS = range(4)
VecSimMetric_L2, VecSimMetric_IP, VecSimMetric_Cosine = range(3)
BY_SCORE, BY_ID, BY_SCORE_THEN_ID = range(3)
VecSim_QueryReply_OK, VecSim_QueryReply_TimedOut = 0, 1

TYPE_SIZE = {0: 4, 1: 8, 2: 2, 3: 2, 4: 1, 5: 1}


class BFParams(C.Structure):
    _fields_ = [("type", C.c_int), ("dim", C.c_size_t), ("metric", C.c_int), ("multi", C.c_bool),
                ("initialCapacity", C.c_size_t), ("blockSize", C.c_size_t)]


class HNSWParams(C.Structure):
    _fields_ = [("type", C.c_int), ("dim", C.c_size_t), ("metric", C.c_int), ("multi", C.c_bool),
                ("initialCapacity", C.c_size_t), ("blockSize", C.c_size_t), ("M", C.c_size_t),
                ("efConstruction", C.c_size_t), ("efRuntime", C.c_size_t), ("epsilon", C.c_double)]


JOB_CB = C.CFUNCTYPE(None, C.c_void_p)
SUBMIT_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(JOB_CB), C.c_size_t)


class TieredHNSWParams(C.Structure):
    _fields_ = [("swapJobThreshold", C.c_size_t)]


class _TieredSpecific(C.Union):
    _fields_ = [("tieredHnswParams", TieredHNSWParams), ("_opaque", C.c_uint64 * 3)]


class TieredIndexParams(C.Structure):
    _fields_ = [("jobQueue", C.c_void_p), ("jobQueueCtx", C.c_void_p), ("submitCb", SUBMIT_CB),
                ("flatBufferLimit", C.c_size_t), ("primaryIndexParams", C.c_void_p), ("specificParams", _TieredSpecific)]


class AlgoParams(C.Union):
    _fields_ = [("hnswParams", HNSWParams), ("bfParams", BFParams), ("tieredParams", TieredIndexParams),
                ("_opaque", C.c_uint64 * 15)]


class VecSimParams(C.Structure):
    _fields_ = [("algo", C.c_int), ("algoParams", AlgoParams), ("logCtx", C.c_void_p)]


class HNSWRuntimeParams(C.Structure):
    _fields_ = [("efRuntime", C.c_size_t), ("epsilon", C.c_double)]


class _RuntimeUnion(C.Union):
    _fields_ = [("hnswRuntimeParams", HNSWRuntimeParams), ("_opaque", C.c_uint64 * 4)]


class VecSimQueryParams(C.Structure):
    _anonymous_ = ("u",)
    _fields_ = [("u", _RuntimeUnion), ("batchSize", C.c_size_t), ("searchMode", C.c_int),
                ("timeoutCtx", C.c_void_p)]


class VecSimIndexBasicInfo(C.Structure):
    _fields_ = [("algo", C.c_int), ("metric", C.c_int), ("type", C.c_int), ("isMulti", C.c_bool),
                ("isTiered", C.c_bool), ("isDisk", C.c_bool), ("blockSize", C.c_size_t), ("dim", C.c_size_t)]


class VecSimIndexStatsInfo(C.Structure):
    _fields_ = [("memory", C.c_size_t), ("numberOfMarkedDeleted", C.c_size_t),
                ("directHNSWInsertions", C.c_size_t), ("flatBufferSize", C.c_size_t)]


class FieldValue(C.Union):
    _fields_ = [("floatingPointValue", C.c_double), ("integerValue", C.c_int64), ("uintegerValue", C.c_uint64),
                ("stringValue", C.c_char_p), ("iteratorValue", C.c_void_p)]


class VecSim_InfoField(C.Structure):
    _fields_ = [("fieldName", C.c_char_p), ("fieldType", C.c_int), ("fieldValue", FieldValue)]


assert C.sizeof(BFParams) == 40 and C.sizeof(VecSimParams) == 136 and C.sizeof(VecSimQueryParams) == 56
assert C.sizeof(TieredIndexParams) == 64

TIMEOUT_CB = C.CFUNCTYPE(C.c_int, C.c_void_p)

_lib = None

EXPORTS = [
    "VecSimIndex_New", "VecSimIndex_EstimateInitialSize", "VecSimIndex_EstimateElementSize", "VecSimIndex_Free",
    "VecSimIndex_AddVector", "VecSimIndex_DeleteVector", "VecSimIndex_GetDistanceFrom_Unsafe", "VecSim_Normalize",
    "VecSimParams_GetQueryBlobSize", "VecSimIndex_IndexSize", "VecSimIndex_ResolveParams", "VecSimIndex_TopKQuery",
    "VecSimIndex_RangeQuery", "VecSimIndex_DebugInfo", "VecSimIndex_BasicInfo", "VecSimIndex_StatsInfo",
    "VecSimBatchIterator_New", "VecSimIndex_PreferAdHocSearch", "VecSimIndex_AdhocBfCtx_New",
    "VecSimIndex_AdhocBfCtx_Free", "VecSimIndex_AdhocBfCtx_GetDistanceFrom", "VecSimIndex_AdhocBfCtx_GetExactDistances",
    "VecSimTieredIndex_GC", "VecSimTieredIndex_AcquireSharedLocks", "VecSimTieredIndex_ReleaseSharedLocks",
    "VecSim_SetMemoryFunctions", "VecSim_SetTimeoutCallbackFunction", "VecSim_SetLogCallbackFunction",
    "VecSim_SetTestLogContext", "VecSim_SetWriteMode", "VecSim_UpdateThreadPoolSize", "VecSim_GetSharedMemory",
    "VecSimQueryResult_GetId", "VecSimQueryResult_GetScore", "VecSimQueryReply_Len", "VecSimQueryReply_GetCode",
    "VecSimQueryReply_Free", "VecSimQueryReply_GetIterator", "VecSimQueryReply_IteratorNext",
    "VecSimQueryReply_IteratorHasNext", "VecSimQueryReply_IteratorReset", "VecSimQueryReply_IteratorFree",
    "VecSimBatchIterator_Next", "VecSimBatchIterator_HasNext", "VecSimBatchIterator_Free", "VecSimBatchIterator_Reset",
    "VecSimIndex_TopKQueryBatch", "VecSimIndex_TopKQueryBatchRaw", "VecSimIndex_AddVectorBatch",
    "VecSimGPU_SetDevice", "VecSimGPU_GetDevice", "VecSimGPU_DeviceCount", "VecSimGPU_SetTopKMode",
    "VecSimGPU_Configure", "VecSimGPU_ShardCount",
    "VecSimGPU_LastQueryStats", "VecSimGPU_GetStore", "VecSimGPU_LastError", "VecSimGPU_AppendDeviceRows",
    "VecSimGPU_HNSWLoadIndex", "VecSimGPU_HNSWSaveIndex",
    "VecSimGPU_GetGraph", "VecSimGPU_HNSWImportGraph", "VecSimGPU_HNSWExportGraph", "VecSimGPU_HNSWLastStats",
    "VecSimIndex_DebugInfoIterator", "VecSimDebugInfoIterator_NumberOfFields", "VecSimDebugInfoIterator_HasNextField",
    "VecSimDebugInfoIterator_NextField", "VecSimDebugInfoIterator_Free", "VecSimDebug_GetElementNeighborsInHNSWGraph",
    "VecSimDebug_ReleaseElementNeighborsInHNSWGraph",
]


def lib_path():
    return os.path.join(HERE, "libvecsim_b200.so")


def lib():
    """Load libvecsim_b200.so (and with it libvsgpu.so). Raises if the libraries are not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(lib_path()):
        raise RuntimeError("libvecsim_b200.so is not built: run `python -m vectorsimilarity_b200.build` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(lib_path())
    vp, sz, i32, dbl, i64 = C.c_void_p, C.c_size_t, C.c_int, C.c_double, C.c_int64
    L.VecSimIndex_New.restype = vp
    L.VecSimIndex_New.argtypes = [C.POINTER(VecSimParams)]
    L.VecSimIndex_Free.argtypes = [vp]
    L.VecSimIndex_AddVector.argtypes = [vp, vp, sz]
    L.VecSimIndex_DeleteVector.argtypes = [vp, sz]
    L.VecSimIndex_GetDistanceFrom_Unsafe.restype = dbl
    L.VecSimIndex_GetDistanceFrom_Unsafe.argtypes = [vp, sz, vp]
    L.VecSim_Normalize.argtypes = [vp, sz, i32]
    L.VecSimParams_GetQueryBlobSize.restype = sz
    L.VecSimParams_GetQueryBlobSize.argtypes = [i32, sz, i32]
    L.VecSimIndex_IndexSize.restype = sz
    L.VecSimIndex_IndexSize.argtypes = [vp]
    L.VecSimIndex_TopKQuery.restype = vp
    L.VecSimIndex_TopKQuery.argtypes = [vp, vp, sz, C.POINTER(VecSimQueryParams), i32]
    L.VecSimIndex_RangeQuery.restype = vp
    L.VecSimIndex_RangeQuery.argtypes = [vp, vp, dbl, C.POINTER(VecSimQueryParams), i32]
    L.VecSimIndex_BasicInfo.restype = VecSimIndexBasicInfo
    L.VecSimIndex_BasicInfo.argtypes = [vp]
    L.VecSimIndex_StatsInfo.restype = VecSimIndexStatsInfo
    L.VecSimIndex_StatsInfo.argtypes = [vp]
    L.VecSimBatchIterator_New.restype = vp
    L.VecSimBatchIterator_New.argtypes = [vp, vp, C.POINTER(VecSimQueryParams)]
    L.VecSimIndex_PreferAdHocSearch.restype = C.c_bool
    L.VecSimIndex_PreferAdHocSearch.argtypes = [vp, sz, sz, C.c_bool]
    L.VecSimIndex_AdhocBfCtx_New.restype = vp
    L.VecSimIndex_AdhocBfCtx_New.argtypes = [vp, vp]
    L.VecSimIndex_AdhocBfCtx_Free.argtypes = [vp]
    L.VecSimIndex_AdhocBfCtx_GetDistanceFrom.restype = dbl
    L.VecSimIndex_AdhocBfCtx_GetDistanceFrom.argtypes = [vp, sz]
    L.VecSimIndex_AdhocBfCtx_GetExactDistances.argtypes = [vp, vp, vp, sz]
    L.VecSim_SetTimeoutCallbackFunction.argtypes = [TIMEOUT_CB]
    L.VecSimQueryResult_GetId.restype = i64
    L.VecSimQueryResult_GetId.argtypes = [vp]
    L.VecSimQueryResult_GetScore.restype = dbl
    L.VecSimQueryResult_GetScore.argtypes = [vp]
    L.VecSimQueryReply_Len.restype = sz
    L.VecSimQueryReply_Len.argtypes = [vp]
    L.VecSimQueryReply_GetCode.restype = i32
    L.VecSimQueryReply_GetCode.argtypes = [vp]
    L.VecSimQueryReply_Free.argtypes = [vp]
    L.VecSimQueryReply_GetIterator.restype = vp
    L.VecSimQueryReply_GetIterator.argtypes = [vp]
    L.VecSimQueryReply_IteratorNext.restype = vp
    L.VecSimQueryReply_IteratorNext.argtypes = [vp]
    L.VecSimQueryReply_IteratorHasNext.restype = C.c_bool
    L.VecSimQueryReply_IteratorHasNext.argtypes = [vp]
    L.VecSimQueryReply_IteratorFree.argtypes = [vp]
    L.VecSimBatchIterator_Next.restype = vp
    L.VecSimBatchIterator_Next.argtypes = [vp, sz, i32]
    L.VecSimBatchIterator_HasNext.restype = C.c_bool
    L.VecSimBatchIterator_HasNext.argtypes = [vp]
    L.VecSimBatchIterator_Free.argtypes = [vp]
    L.VecSimBatchIterator_Reset.argtypes = [vp]
    L.VecSimIndex_TopKQueryBatch.argtypes = [vp, vp, sz, sz, C.POINTER(VecSimQueryParams), i32, C.POINTER(vp)]
    L.VecSimIndex_TopKQueryBatchRaw.argtypes = [vp, vp, sz, sz, C.POINTER(VecSimQueryParams), vp, vp]
    L.VecSimIndex_AddVectorBatch.restype = C.c_long
    L.VecSimIndex_AddVectorBatch.argtypes = [vp, vp, sz, vp, sz]
    L.VecSimGPU_SetDevice.argtypes = [i32]
    L.VecSimGPU_SetTopKMode.argtypes = [i32]
    L.VecSimGPU_LastQueryStats.argtypes = [vp, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.VecSimGPU_AppendDeviceRows.restype = C.c_long
    L.VecSimGPU_AppendDeviceRows.argtypes = [vp, vp, sz, sz, sz]
    L.VecSimGPU_GetStore.restype = vp
    L.VecSimGPU_GetStore.argtypes = [vp]
    L.VecSimGPU_LastError.restype = C.c_char_p
    L.VecSimIndex_DebugInfoIterator.restype = vp
    L.VecSimIndex_DebugInfoIterator.argtypes = [vp]
    L.VecSimDebugInfoIterator_NumberOfFields.restype = sz
    L.VecSimDebugInfoIterator_NumberOfFields.argtypes = [vp]
    L.VecSimDebugInfoIterator_HasNextField.restype = C.c_bool
    L.VecSimDebugInfoIterator_HasNextField.argtypes = [vp]
    L.VecSimDebugInfoIterator_NextField.restype = C.POINTER(VecSim_InfoField)
    L.VecSimDebugInfoIterator_NextField.argtypes = [vp]
    L.VecSimDebugInfoIterator_Free.argtypes = [vp]
    L.VecSimDebug_GetElementNeighborsInHNSWGraph.argtypes = [vp, sz, C.POINTER(C.POINTER(C.POINTER(C.c_int)))]
    L.VecSimDebug_ReleaseElementNeighborsInHNSWGraph.argtypes = [C.POINTER(C.POINTER(C.c_int))]
    L.VecSimGPU_GetGraph.restype = vp
    L.VecSimGPU_GetGraph.argtypes = [vp]
    L.VecSimGPU_HNSWImportGraph.argtypes = [vp, vp, i32, sz, vp, vp, vp, vp, sz, C.c_long, C.c_long]
    L.VecSimGPU_HNSWExportGraph.argtypes = [vp, vp, vp, vp, sz, C.POINTER(sz), C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.VecSimGPU_HNSWLastStats.argtypes = [vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_float)]
    L.VecSimGPU_HNSWLoadIndex.restype = vp
    L.VecSimGPU_HNSWLoadIndex.argtypes = [C.c_char_p]
    L.VecSimGPU_HNSWSaveIndex.argtypes = [vp, C.c_char_p]
    L.VecSim_SetWriteMode.argtypes = [i32]
    L.VecSimTieredIndex_GC.argtypes = [vp]
    L.VecSimTieredIndex_AcquireSharedLocks.argtypes = [vp]
    L.VecSimTieredIndex_ReleaseSharedLocks.argtypes = [vp]
    L.VecSimGPU_Configure.argtypes = [C.POINTER(C.c_int), C.c_size_t]
    L.VecSimGPU_ShardCount.restype = C.c_size_t
    L.VecSimGPU_ShardCount.argtypes = [vp]
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _drain(rep):
    """VecSimQueryReply* -> (labels, scores, code) through the reply-iterator API; frees the reply."""
    L = lib()
    n = L.VecSimQueryReply_Len(rep)
    code = L.VecSimQueryReply_GetCode(rep)
    labels = np.empty(n, dtype=np.int64)
    scores = np.empty(n, dtype=np.float64)
    it = L.VecSimQueryReply_GetIterator(rep)
    i = 0
    while L.VecSimQueryReply_IteratorHasNext(it):
        item = L.VecSimQueryReply_IteratorNext(it)
        labels[i] = L.VecSimQueryResult_GetId(item)
        scores[i] = L.VecSimQueryResult_GetScore(item)
        i += 1
    assert i == n
    L.VecSimQueryReply_IteratorFree(it)
    L.VecSimQueryReply_Free(rep)
    return labels, scores, code


_timeout_keepalive = None


def set_timeout_callback(fn):
    """fn(ctx) -> int, or None to clear (VecSim_SetTimeoutCallbackFunction)."""
    global _timeout_keepalive
    cb = TIMEOUT_CB(fn) if fn is not None else TIMEOUT_CB()
    _timeout_keepalive = cb
    lib().VecSim_SetTimeoutCallbackFunction(cb)


def set_topk_mode(mode):
    lib().VecSimGPU_SetTopKMode(mode)


def device_count():
    return lib().VecSimGPU_DeviceCount()


def set_device(d):
    if lib().VecSimGPU_SetDevice(d) != 0:
        raise RuntimeError("no such CUDA device %d" % d)


def configure_devices(devices):
    """VecSimGPU_Configure: flat indexes created afterwards shard their rows over `devices` (one process, all GPUs)."""
    arr = (C.c_int * len(devices))(*devices)
    if lib().VecSimGPU_Configure(arr, len(devices)) != 0:
        raise RuntimeError("VecSimGPU_Configure: bad device list %r" % (devices,))


def normalize(blob, dim, vtype):
    lib().VecSim_Normalize(_ptr(blob), dim, vtype)
    return blob


class VecSimIndex:
    def __init__(self, params):
        self._h = lib().VecSimIndex_New(C.byref(params))
        if not self._h:
            raise RuntimeError("VecSimIndex_New returned NULL: " + (lib().VecSimGPU_LastError() or b"").decode())
        info = lib().VecSimIndex_BasicInfo(self._h)
        self.type, self.dim, self.metric = info.type, info.dim, info.metric
        self._blob = TYPE_SIZE[self.type] * self.dim

    def close(self):
        if getattr(self, "_h", None):
            lib().VecSimIndex_Free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, blob):
        blob = np.ascontiguousarray(blob)
        assert blob.nbytes >= self._blob, "query/vector blob too small"
        return blob

    def add_vector(self, blob, label):
        blob = self._check(blob)
        return lib().VecSimIndex_AddVector(self._h, _ptr(blob), int(label))

    def add_vectors(self, blobs, labels=None, first_label=0):
        blobs = np.ascontiguousarray(blobs)
        assert blobs.ndim == 2 and blobs.shape[1] * blobs.itemsize == self._blob
        lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.uint64)
        n = lib().VecSimIndex_AddVectorBatch(self._h, _ptr(blobs), blobs.shape[0], _ptr(lab) if lab is not None else None,
                                             first_label)
        if n < 0:
            raise RuntimeError("AddVectorBatch failed: " + lib().VecSimGPU_LastError().decode())
        return n

    def add_device_rows(self, device_ptr, stride_bytes, n, first_label):
        """Bulk ingest of processed rows already resident on the index's GPU (VecSimGPU_AppendDeviceRows)."""
        r = lib().VecSimGPU_AppendDeviceRows(self._h, C.c_void_p(device_ptr), stride_bytes, n, first_label)
        if r < 0:
            raise RuntimeError("AppendDeviceRows failed: " + lib().VecSimGPU_LastError().decode())
        return r

    def delete_vector(self, label):
        return lib().VecSimIndex_DeleteVector(self._h, int(label))

    def index_size(self):
        return lib().VecSimIndex_IndexSize(self._h)

    def index_memory(self):
        return lib().VecSimIndex_StatsInfo(self._h).memory

    def get_distance_from(self, label, blob):
        blob = np.ascontiguousarray(blob)
        return lib().VecSimIndex_GetDistanceFrom_Unsafe(self._h, int(label), _ptr(blob))

    def knn_query(self, vector, k, query_param=None, order=BY_SCORE):
        """-> (labels[1,n], distances[1,n]) like the reference binding; .last_code holds the reply code."""
        vector = self._check(vector)
        rep = lib().VecSimIndex_TopKQuery(self._h, _ptr(vector), k, C.byref(query_param) if query_param else None, order)
        labels, scores, code = _drain(rep)
        self.last_code = code
        return labels.reshape(1, -1), scores.reshape(1, -1)

    def range_query(self, vector, radius, query_param=None, order=BY_SCORE):
        vector = self._check(vector)
        rep = lib().VecSimIndex_RangeQuery(self._h, _ptr(vector), float(radius),
                                           C.byref(query_param) if query_param else None, order)
        labels, scores, code = _drain(rep)
        self.last_code = code
        return labels.reshape(1, -1), scores.reshape(1, -1)

    def knn_batch(self, queries, k, query_param=None):
        """Batched extension (VecSimIndex_TopKQueryBatchRaw): -> labels[nq,k] (int64, -1 padded), scores[nq,k]."""
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        assert queries.shape[1] * queries.itemsize == self._blob
        labels = np.empty((nq, k), dtype=np.uint64)
        scores = np.empty((nq, k), dtype=np.float64)
        rc = lib().VecSimIndex_TopKQueryBatchRaw(self._h, _ptr(queries), nq, k,
                                                 C.byref(query_param) if query_param else None, _ptr(labels), _ptr(scores))
        if rc < 0:
            raise RuntimeError("TopKQueryBatchRaw failed: " + lib().VecSimGPU_LastError().decode())
        self.last_code = rc
        return labels.view(np.int64), scores

    def knn_batch_replies(self, queries, k, order=BY_SCORE, query_param=None):
        """VecSimIndex_TopKQueryBatch: one reply object per query, drained through the reply API."""
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        out = (C.c_void_p * nq)()
        rc = lib().VecSimIndex_TopKQueryBatch(self._h, _ptr(queries), nq, k, C.byref(query_param) if query_param else None,
                                              order, out)
        if rc != 0:
            raise RuntimeError("TopKQueryBatch failed: " + lib().VecSimGPU_LastError().decode())
        return [_drain(out[i]) for i in range(nq)]

    def create_batch_iterator(self, query_blob, query_param=None):
        query_blob = self._check(query_blob)
        return BatchIterator(self, lib().VecSimBatchIterator_New(self._h, _ptr(query_blob),
                                                                 C.byref(query_param) if query_param else None))

    def prefer_adhoc(self, subset_size, k, initial_check=True):
        return lib().VecSimIndex_PreferAdHocSearch(self._h, subset_size, k, initial_check)

    def adhoc_distances(self, query_blob, labels):
        query_blob = self._check(query_blob)
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        out = np.empty(labels.size, dtype=np.float64)
        ctx = lib().VecSimIndex_AdhocBfCtx_New(self._h, _ptr(query_blob))
        lib().VecSimIndex_AdhocBfCtx_GetExactDistances(ctx, _ptr(labels), _ptr(out), labels.size)
        lib().VecSimIndex_AdhocBfCtx_Free(ctx)
        return out

    def last_query_stats(self):
        path, launches, fb = C.c_uint(), C.c_uint(), C.c_uint()
        cand = C.c_uint64()
        scan, total = C.c_float(), C.c_float()
        lib().VecSimGPU_LastQueryStats(self._h, C.byref(path), C.byref(launches), C.byref(cand), C.byref(fb),
                                       C.byref(scan), C.byref(total))
        return dict(path=path.value, kernel_launches=launches.value, candidates=cand.value, fallback_queries=fb.value,
                    scan_ms=scan.value, total_ms=total.value)

    def shard_count(self):
        return lib().VecSimGPU_ShardCount(self._h)

    def device_store(self):
        return lib().VecSimGPU_GetStore(self._h)

    def debug_info(self):
        """VecSimIndex_DebugInfoIterator drained into an ordered list of (name, value)."""
        return self._drain_info(lib().VecSimIndex_DebugInfoIterator(self._h), free=True)

    def _drain_info(self, it, free):
        L = lib()
        out = []
        n = L.VecSimDebugInfoIterator_NumberOfFields(it)
        while L.VecSimDebugInfoIterator_HasNextField(it):
            f = L.VecSimDebugInfoIterator_NextField(it).contents
            # read only the union member the type names (a char* view of an integer field would be dereferenced)
            if f.fieldType == 0:
                v = f.fieldValue.stringValue
            elif f.fieldType == 1:
                v = f.fieldValue.integerValue
            elif f.fieldType == 2:
                v = f.fieldValue.uintegerValue
            elif f.fieldType == 4:  # nested iterator (tiered: FRONTEND_INDEX / BACKEND_INDEX), owned by its parent
                v = self._drain_info(f.fieldValue.iteratorValue, free=False)
            else:
                v = f.fieldValue.floatingPointValue
            out.append((f.fieldName.decode(), v.decode() if isinstance(v, bytes) else v))
        assert len(out) == n
        if free:
            L.VecSimDebugInfoIterator_Free(it)
        return out


class BFIndex(VecSimIndex):
    def __init__(self, params):
        p = VecSimParams()
        p.algo = VecSimAlgo_BF
        p.algoParams.bfParams = params
        super().__init__(p)


class HNSWIndex(VecSimIndex):
    """Mirror of the reference binding's HNSWIndex (src/python_bindings/bindings.cpp:286-420): add_vector,
    knn_query, range_query, set_ef; plus the bulk graph import / export of include/vecsim_b200.h."""

    def __init__(self, params):
        p = VecSimParams()
        p.algo = VecSimAlgo_HNSWLIB
        p.algoParams.hnswParams = params
        super().__init__(p)
        self.M = params.M or 16
        self._ef = 0

    @classmethod
    def load(cls, path):
        """HNSWIndex(file_name) of the reference binding (bindings.cpp:300-304): an index restored from a serialized file."""
        h = lib().VecSimGPU_HNSWLoadIndex(os.fsencode(path))
        if not h:
            raise RuntimeError("VecSimGPU_HNSWLoadIndex: " + (lib().VecSimGPU_LastError() or b"").decode())
        self = cls.__new__(cls)
        self._h = h
        info = lib().VecSimIndex_BasicInfo(h)
        self.type, self.dim, self.metric = info.type, info.dim, info.metric
        self._blob = TYPE_SIZE[self.type] * self.dim
        self.M = dict(self.debug_info())["M"]
        self._ef = 0
        return self

    def save_index(self, path):
        if lib().VecSimGPU_HNSWSaveIndex(self._h, os.fsencode(path)) != 0:
            raise RuntimeError("VecSimGPU_HNSWSaveIndex failed: " + (lib().VecSimGPU_LastError() or b"").decode())

    def set_ef(self, ef):
        self._ef = int(ef)

    def _qp(self, query_param):
        if query_param is not None or not self._ef:
            return query_param
        qp = VecSimQueryParams()
        qp.hnswRuntimeParams.efRuntime = self._ef
        return qp

    def knn_query(self, vector, k, query_param=None, order=BY_SCORE):
        return super().knn_query(vector, k, self._qp(query_param), order)

    def knn_batch(self, queries, k, query_param=None):
        return super().knn_batch(queries, k, self._qp(query_param))

    def import_graph(self, vectors, levels, links, counts, entry, max_level, labels=None, processed=True):
        """Adopt a graph over `vectors` (rows in internal-id order). links[l]: [n, width_l] u32, counts[l]: [n]
        (the layout oracle.ref.RefIndex.hnsw_export produces)."""
        vectors = np.ascontiguousarray(vectors)
        n = vectors.shape[0]
        levels = np.ascontiguousarray(levels, dtype=np.uint32)
        M, M0 = self.M, 2 * self.M
        l0 = np.zeros((n, M0 + 1), dtype=np.uint32)
        l0[:, 0] = counts[0]
        l0[:, 1:] = np.where(np.arange(M0)[None, :] < np.asarray(counts[0])[:, None], links[0], 0)
        recs = []
        for i in np.nonzero(levels)[0]:
            for lvl in range(1, int(levels[i]) + 1):
                r = np.zeros(M + 1, dtype=np.uint32)
                c = int(counts[lvl][i])
                r[0] = c
                r[1:1 + c] = links[lvl][i][:c]
                recs.append(r)
        upper = np.ascontiguousarray(np.stack(recs)) if recs else np.zeros((0, M + 1), dtype=np.uint32)
        lab = None if labels is None else np.ascontiguousarray(labels, dtype=np.uint64)
        rc = lib().VecSimGPU_HNSWImportGraph(self._h, _ptr(vectors), int(processed), n, _ptr(lab) if lab is not None else None,
                                             _ptr(levels), _ptr(l0), _ptr(upper) if len(recs) else None, len(recs),
                                             int(entry), int(max_level))
        if rc != 0:
            raise RuntimeError("HNSWImportGraph failed: " + lib().VecSimGPU_LastError().decode())

    def export_graph(self, n):
        """-> dict(levels[n], l0[n, 2M+1], upper[records, M+1], entry, max_level)."""
        M, M0 = self.M, 2 * self.M
        levels = np.zeros(n, dtype=np.uint32)
        l0 = np.zeros((n, M0 + 1), dtype=np.uint32)
        recs = C.c_size_t()
        entry, maxl = C.c_long(), C.c_long()
        rc = lib().VecSimGPU_HNSWExportGraph(self._h, _ptr(levels), _ptr(l0), None, 0, C.byref(recs), C.byref(entry),
                                             C.byref(maxl))
        assert rc == 0, lib().VecSimGPU_LastError()
        upper = np.zeros((recs.value, M + 1), dtype=np.uint32)
        if recs.value:
            rc = lib().VecSimGPU_HNSWExportGraph(self._h, None, None, _ptr(upper), recs.value, C.byref(recs),
                                                 C.byref(entry), C.byref(maxl))
            assert rc == 0, lib().VecSimGPU_LastError()
        return dict(levels=levels, l0=l0, upper=upper, entry=entry.value, max_level=maxl.value)

    def element_neighbors(self, label):
        """VecSimDebug_GetElementNeighborsInHNSWGraph -> list (one per level) of neighbour labels."""
        L = lib()
        out = C.POINTER(C.POINTER(C.c_int))()
        rc = L.VecSimDebug_GetElementNeighborsInHNSWGraph(self._h, int(label), C.byref(out))
        if rc != 0:
            return rc, None
        levels = []
        i = 0
        while out[i]:
            cnt = out[i][0]
            levels.append([out[i][1 + j] for j in range(cnt)])
            i += 1
        L.VecSimDebug_ReleaseElementNeighborsInHNSWGraph(out)
        return 0, levels

    def hnsw_stats(self):
        ev, hops, ms = C.c_ulonglong(), C.c_ulonglong(), C.c_float()
        lib().VecSimGPU_HNSWLastStats(self._h, C.byref(ev), C.byref(hops), C.byref(ms))
        return dict(dist_evals=ev.value, hops=hops.value, ms=ms.value)


class Tiered_HNSWIndex(HNSWIndex):
    """Mirror of the reference binding's Tiered_HNSWIndex (src/python_bindings/bindings.cpp:486-560): the index is
    created over a mock job queue owned by this object; add_vector files an insert job, wait_for_index runs the
    queued jobs (on `threads` Python threads, as RediSearch's workers would) until the flat buffer is drained."""

    def __init__(self, hnsw_params, tiered_hnsw_params=None, flat_buffer_size=1024, threads=0):
        import collections
        import threading
        self._jobs = collections.deque()
        self._jobs_lock = threading.Lock()
        self._threads = threads

        def submit(_queue, _ctx, jobs, cbs, n):
            with self._jobs_lock:
                for i in range(n):
                    self._jobs.append((jobs[i], C.cast(cbs[i], C.c_void_p).value))
            return 0

        self._submit = SUBMIT_CB(submit)  # kept alive with the index
        self._primary = VecSimParams()
        self._primary.algo = VecSimAlgo_HNSWLIB
        self._primary.algoParams.hnswParams = hnsw_params
        p = VecSimParams()
        p.algo = VecSimAlgo_TIERED
        tp = p.algoParams.tieredParams
        tp.jobQueue = None
        tp.jobQueueCtx = None
        tp.submitCb = self._submit
        tp.flatBufferLimit = flat_buffer_size
        tp.primaryIndexParams = C.cast(C.pointer(self._primary), C.c_void_p)
        tp.specificParams.tieredHnswParams.swapJobThreshold = tiered_hnsw_params.swapJobThreshold if tiered_hnsw_params else 0
        VecSimIndex.__init__(self, p)
        self.M = hnsw_params.M or 16
        self._ef = 0
        self._buffer_limit = flat_buffer_size

    def pending_jobs(self):
        with self._jobs_lock:
            return len(self._jobs)

    def run_jobs(self, max_jobs=None):
        """Execute queued jobs on the calling thread; -> number run."""
        done = 0
        while max_jobs is None or done < max_jobs:
            with self._jobs_lock:
                if not self._jobs:
                    break
                job, cb = self._jobs.popleft()
            JOB_CB(cb)(job)
            done += 1
        return done

    def wait_for_index(self, waiting_duration=10):
        if self._threads <= 1:
            self.run_jobs()
            return
        import threading
        ts = [threading.Thread(target=self.run_jobs) for _ in range(self._threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(waiting_duration)

    def get_curr_bf_size(self):
        return lib().VecSimIndex_StatsInfo(self._h).flatBufferSize

    def get_buffer_limit(self):
        return self._buffer_limit

    def get_threads_num(self):
        return max(self._threads, 1)

    def hnsw_label_count(self):
        return self.index_size() - self.get_curr_bf_size()

    def stats(self):
        st = lib().VecSimIndex_StatsInfo(self._h)
        return dict(memory=st.memory, numberOfMarkedDeleted=st.numberOfMarkedDeleted,
                    directHNSWInsertions=st.directHNSWInsertions, flatBufferSize=st.flatBufferSize)

    def close(self):
        # queued jobs that never ran still reference the index: run them (they are cheap no-ops once it is drained)
        if getattr(self, "_h", None):
            self.run_jobs()
        super().close()


def set_write_mode(in_place):
    """VecSim_SetWriteMode: False = VecSim_WriteAsync (default), True = VecSim_WriteInPlace."""
    lib().VecSim_SetWriteMode(1 if in_place else 0)


class BatchIterator:
    def __init__(self, index, handle):
        self._index, self._h = index, handle

    def has_next(self):
        return lib().VecSimBatchIterator_HasNext(self._h)

    def get_next_results(self, n, order=BY_SCORE):
        labels, scores, code = _drain(lib().VecSimBatchIterator_Next(self._h, n, order))
        self.last_code = code
        return labels.reshape(1, -1), scores.reshape(1, -1)

    def reset(self):
        lib().VecSimBatchIterator_Reset(self._h)

    def close(self):
        if self._h:
            lib().VecSimBatchIterator_Free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
