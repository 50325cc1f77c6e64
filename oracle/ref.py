"""oracle/ref.py — TEST INFRASTRUCTURE ONLY: ctypes view of oracle/_ref/libvecsim_ref.so.

The .so is the UNMODIFIED reference (RedisAI/VectorSimilarity) compiled by oracle/Makefile from
/root/reference plus the C façade in oracle/ref_harness.cpp. Only tests/, bench.py's
cpu_baseline / --impl reference legs and __graft_entry__.smoke() may import this module.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libvecsim_ref.so")

FLOAT32, FLOAT64, BFLOAT16, FLOAT16, INT8, UINT8 = range(6)
L2, IP, COSINE = range(3)
BY_SCORE, BY_ID = 0, 1

FEATURE_BITS = {n: 1 << i for i, n in enumerate(
    ["sse", "sse3", "sse4_1", "avx", "avx2", "fma3", "f16c", "avx512f", "avx512bw", "avx512vl",
     "avx512vnni", "avx512vbmi2", "avx512_bf16", "avx512_fp16"])}

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, sz, i32, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.vsref_set_feature_disable_mask.argtypes = [C.c_uint32]
        L.vsref_host_features.restype = C.c_uint32
        L.vsref_set_timeout.argtypes = [i32]
        L.vsref_distance.argtypes = [i32, i32, sz, vp, vp, C.POINTER(dbl)]
        L.vsref_distance_many.argtypes = [i32, i32, sz, vp, sz, vp, sz, sz, vp]
        L.vsref_normalize.argtypes = [i32, sz, vp]
        L.vsref_bf_new.restype = vp
        L.vsref_bf_new.argtypes = [i32, sz, i32, i32, sz]
        L.vsref_hnsw_new.restype = vp
        L.vsref_hnsw_new.argtypes = [i32, sz, i32, i32, sz, sz, sz, sz]
        L.vsref_index_free.argtypes = [vp]
        L.vsref_add.argtypes = [vp, vp, sz]
        L.vsref_add_many.restype = C.c_long
        L.vsref_add_many.argtypes = [vp, vp, sz, sz, vp, sz]
        L.vsref_delete.argtypes = [vp, sz]
        L.vsref_size.restype = sz
        L.vsref_size.argtypes = [vp]
        L.vsref_label_count.restype = sz
        L.vsref_label_count.argtypes = [vp]
        L.vsref_distance_from.restype = dbl
        L.vsref_distance_from.argtypes = [vp, sz, vp]
        L.vsref_topk.restype = sz
        L.vsref_topk.argtypes = [vp, vp, sz, i32, sz, vp, vp, C.POINTER(i32)]
        L.vsref_range.restype = C.c_long
        L.vsref_range.argtypes = [vp, vp, dbl, i32, sz, vp, vp, C.POINTER(i32)]
        L.vsref_bi_new.restype = vp
        L.vsref_bi_new.argtypes = [vp, vp]
        L.vsref_bi_next.restype = sz
        L.vsref_bi_next.argtypes = [vp, sz, i32, vp, vp, C.POINTER(i32)]
        L.vsref_bi_has_next.argtypes = [vp]
        L.vsref_bi_reset.argtypes = [vp]
        L.vsref_bi_free.argtypes = [vp]
        L.vsref_topk_many.restype = dbl
        L.vsref_topk_many.argtypes = [vp, vp, sz, sz, sz, sz, i32, vp, vp]
        L.vsref_hnsw_info.argtypes = [vp] + [C.POINTER(sz)] * 3 + [C.POINTER(C.c_long)] * 2
        L.vsref_hnsw_export_meta.argtypes = [vp, vp, vp, vp]
        L.vsref_hnsw_export_level.argtypes = [vp, sz, sz, vp, vp]
        L.vsref_hnsw_export_vectors.argtypes = [vp, sz, sz, vp]
        L.vsref_silence_logs()
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def set_disabled_features(*names):
    m = 0
    for n in names:
        m |= FEATURE_BITS[n]
    lib().vsref_set_feature_disable_mask(m)


def host_features():
    m = lib().vsref_host_features()
    return [n for n, b in FEATURE_BITS.items() if m & b]


def set_timeout(flag):
    lib().vsref_set_timeout(int(flag))


def distance(vtype, metric, a, b, dim=None):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    out = C.c_double()
    if dim is None:
        dim = a.size
    rc = lib().vsref_distance(vtype, metric, dim, _ptr(a), _ptr(b), C.byref(out))
    assert rc == 0
    return out.value


def distance_many(vtype, metric, dim, A, B):
    """A, B: 2-D contiguous arrays with n rows each (row = one stored blob)."""
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B)
    n = A.shape[0]
    out = np.empty(n, dtype=np.float64)
    rc = lib().vsref_distance_many(vtype, metric, dim, _ptr(A), A.strides[0], _ptr(B), B.strides[0],
                                   n, _ptr(out))
    assert rc == 0
    return out


def normalize(vtype, dim, blob):
    """In-place. int8/uint8 blobs must have dim+4 bytes."""
    assert lib().vsref_normalize(vtype, dim, _ptr(blob)) == 0
    return blob


class RefIndex:
    """Reference flat (BruteForce) or HNSW index."""

    def __init__(self, vtype, dim, metric, multi=False, block_size=1024, algo="flat", M=16,
                 ef_construction=200, ef_runtime=10):
        L = lib()
        self.vtype, self.dim, self.metric, self.algo = vtype, dim, metric, algo
        if algo == "flat":
            self.h = L.vsref_bf_new(vtype, dim, metric, int(multi), block_size)
        else:
            self.h = L.vsref_hnsw_new(vtype, dim, metric, int(multi), block_size, M, ef_construction,
                                      ef_runtime)
        if not self.h:
            raise RuntimeError("reference index creation failed")

    def close(self):
        if self.h:
            lib().vsref_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, blob, label):
        blob = np.ascontiguousarray(blob)
        return lib().vsref_add(self.h, _ptr(blob), label)

    def add_many(self, blobs, labels=None, first_label=0):
        blobs = np.ascontiguousarray(blobs)
        lab = None
        if labels is not None:
            lab = np.ascontiguousarray(labels, dtype=np.uint64)
        return lib().vsref_add_many(self.h, _ptr(blobs), blobs.strides[0], blobs.shape[0],
                                    _ptr(lab) if lab is not None else None, first_label)

    def delete(self, label):
        return lib().vsref_delete(self.h, label)

    def size(self):
        return lib().vsref_size(self.h)

    def distance_from(self, label, blob):
        blob = np.ascontiguousarray(blob)
        return lib().vsref_distance_from(self.h, label, _ptr(blob))

    def topk(self, q, k, order=BY_SCORE, ef_runtime=0):
        q = np.ascontiguousarray(q)
        labels = np.empty(max(k, 1), dtype=np.uint64)
        scores = np.empty(max(k, 1), dtype=np.float64)
        code = C.c_int()
        n = lib().vsref_topk(self.h, _ptr(q), k, order, ef_runtime, _ptr(labels), _ptr(scores),
                             C.byref(code))
        return labels[:n].copy(), scores[:n].copy(), code.value

    def range(self, q, radius, order=BY_SCORE, cap=None):
        q = np.ascontiguousarray(q)
        cap = cap or max(self.size(), 1)
        labels = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        code = C.c_int()
        n = lib().vsref_range(self.h, _ptr(q), float(radius), order, cap, _ptr(labels), _ptr(scores),
                              C.byref(code))
        if n < 0:
            raise RuntimeError("reference rangeQuery threw")
        return labels[:n].copy(), scores[:n].copy(), code.value

    def topk_many(self, queries, k, n_threads=1, ef_runtime=0, want_results=True):
        queries = np.ascontiguousarray(queries)
        nq = queries.shape[0]
        labels = np.empty((nq, k), dtype=np.uint64) if want_results else None
        scores = np.empty((nq, k), dtype=np.float64) if want_results else None
        secs = lib().vsref_topk_many(self.h, _ptr(queries), queries.strides[0], nq, k, ef_runtime,
                                     n_threads, _ptr(labels) if want_results else None,
                                     _ptr(scores) if want_results else None)
        return labels, scores, secs

    def batch_iterator(self, q):
        return RefBatchIterator(self, q)

    # --- HNSW graph export (fp32 only) ---
    def hnsw_export(self):
        L = lib()
        n, M, ef = C.c_size_t(), C.c_size_t(), C.c_size_t()
        entry, maxl = C.c_long(), C.c_long()
        assert L.vsref_hnsw_info(self.h, C.byref(n), C.byref(M), C.byref(ef), C.byref(entry),
                                 C.byref(maxl)) == 0
        n, M = n.value, M.value
        levels = np.empty(n, dtype=np.uint32)
        labels = np.empty(n, dtype=np.uint64)
        flags = np.empty(n, dtype=np.uint8)
        assert L.vsref_hnsw_export_meta(self.h, _ptr(levels), _ptr(labels), _ptr(flags)) == 0
        out = dict(n=n, M=M, ef=ef.value, entry=entry.value, max_level=maxl.value, levels=levels,
                   labels=labels, flags=flags, links=[], counts=[])
        for lvl in range(max(maxl.value, 0) + 1):
            width = 2 * M if lvl == 0 else M
            links = np.empty((n, width), dtype=np.uint32)
            counts = np.empty(n, dtype=np.uint32)
            assert L.vsref_hnsw_export_level(self.h, lvl, width, _ptr(links), _ptr(counts)) == 0
            out["links"].append(links)
            out["counts"].append(counts)
        vec = np.empty((n, self.dim), dtype=np.float32)
        assert L.vsref_hnsw_export_vectors(self.h, vec.strides[0], self.dim * 4, _ptr(vec)) == 0
        out["vectors"] = vec
        return out


class RefBatchIterator:
    def __init__(self, index, q):
        q = np.ascontiguousarray(q)
        self.index = index
        self.it = lib().vsref_bi_new(index.h, _ptr(q))

    def next(self, n, order=BY_SCORE):
        labels = np.empty(max(n, 1), dtype=np.uint64)
        scores = np.empty(max(n, 1), dtype=np.float64)
        code = C.c_int()
        m = lib().vsref_bi_next(self.it, n, order, _ptr(labels), _ptr(scores), C.byref(code))
        return labels[:m].copy(), scores[:m].copy(), code.value

    def has_next(self):
        return bool(lib().vsref_bi_has_next(self.it))

    def reset(self):
        lib().vsref_bi_reset(self.it)

    def close(self):
        if self.it:
            lib().vsref_bi_free(self.it)
            self.it = None


# ---- BUILD_TESTS variant (make -C oracle ref_bt): the reference's own index-file loader / writer --------------------------
BT_LIB_PATH = os.path.join(_HERE, "_ref", "libvecsim_ref_bt.so")
_bt = None


def bt_available():
    return os.path.exists(BT_LIB_PATH)


def bt_lib():
    global _bt
    if _bt is None:
        L = C.CDLL(BT_LIB_PATH)
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L.vsref_hnsw_load.restype = vp
        L.vsref_hnsw_load.argtypes = [C.c_char_p]
        L.vsref_hnsw_save.argtypes = [vp, C.c_char_p]
        L.vsref_hnsw_integrity.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]
        L.vsref_index_free.argtypes = [vp]
        L.vsref_size.restype = sz
        L.vsref_size.argtypes = [vp]
        L.vsref_topk.restype = sz
        L.vsref_topk.argtypes = [vp, vp, sz, i32, sz, vp, vp, C.POINTER(i32)]
        L.vsref_delete.argtypes = [vp, sz]
        L.vsref_silence_logs()
        _bt = L
    return _bt


class RefFileIndex:
    """An HNSW index restored by the reference's own loader, HNSWFactory::NewIndex(location)
    (index_factories/hnsw_factory.cpp:171-251) — compiled only under BUILD_TESTS, hence the separate library."""

    def __init__(self, path):
        self.h = bt_lib().vsref_hnsw_load(os.fsencode(path))
        if not self.h:
            raise RuntimeError("the reference could not load " + str(path))

    def close(self):
        if self.h:
            bt_lib().vsref_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def size(self):
        return bt_lib().vsref_size(self.h)

    def integrity(self):
        """checkIntegrity (hnsw_serializer_impl.h:54-143, fp32 indexes): (valid, bidirectional, unidirectional)."""
        d, u = C.c_size_t(), C.c_size_t()
        ok = bt_lib().vsref_hnsw_integrity(self.h, C.byref(d), C.byref(u))
        return ok, d.value, u.value

    def save(self, path):
        if bt_lib().vsref_hnsw_save(self.h, os.fsencode(path)) != 0:
            raise RuntimeError("the reference could not save the index")

    def delete(self, label):
        return bt_lib().vsref_delete(self.h, label)

    def topk(self, q, k, ef_runtime=0):
        q = np.ascontiguousarray(q)
        labels = np.empty(max(k, 1), dtype=np.uint64)
        scores = np.empty(max(k, 1), dtype=np.float64)
        code = C.c_int()
        n = bt_lib().vsref_topk(self.h, _ptr(q), k, BY_SCORE, ef_runtime, _ptr(labels), _ptr(scores), C.byref(code))
        return labels[:n].copy(), scores[:n].copy(), code.value
