"""oracle/port.py — TEST INFRASTRUCTURE ONLY: ctypes view of oracle/libvs_oracle.so.

libvs_oracle.so is the plain-C restatement (oracle/vs_oracle.c) of the reference's CPU algorithm for
the flat top-K / range / batch-iterator path. Only tests/, bench.py's cpu_baseline leg and
__graft_entry__.smoke() may import this module; the product path never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvs_oracle.so")

FLOAT32, FLOAT64, BFLOAT16, FLOAT16, INT8, UINT8 = range(6)
L2, IP, COSINE = range(3)
BY_SCORE, BY_ID = 0, 1
TIER_AVX512, TIER_AVX512_NOBF16, TIER_NAIVE = range(3)

NP_DTYPE = {FLOAT32: np.float32, FLOAT64: np.float64, BFLOAT16: np.uint16, FLOAT16: np.float16,
            INT8: np.int8, UINT8: np.uint8}

_lib = None


def build():
    """Compile the C restatement (gcc only; no reference sources involved)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp, sz, i32, dbl = C.c_void_p, C.c_size_t, C.c_int, C.c_double
        L.vso_set_tier.argtypes = [i32]
        L.vso_stored_size.restype = sz
        L.vso_stored_size.argtypes = [i32, i32, sz]
        L.vso_distance.restype = dbl
        L.vso_distance.argtypes = [i32, i32, sz, vp, vp]
        L.vso_normalize.argtypes = [i32, sz, vp]
        L.vso_f32_to_bf16.restype = C.c_uint16
        L.vso_f32_to_bf16.argtypes = [C.c_float]
        L.vso_bf16_to_f32.restype = C.c_float
        L.vso_bf16_to_f32.argtypes = [C.c_uint16]
        L.vso_f32_to_fp16.restype = C.c_uint16
        L.vso_f32_to_fp16.argtypes = [C.c_float]
        L.vso_fp16_to_f32.restype = C.c_float
        L.vso_fp16_to_f32.argtypes = [C.c_uint16]
        L.vso_flat_new.restype = vp
        L.vso_flat_new.argtypes = [i32, sz, i32, i32, sz]
        L.vso_flat_free.argtypes = [vp]
        L.vso_flat_add.argtypes = [vp, vp, sz]
        L.vso_flat_delete.argtypes = [vp, sz]
        L.vso_flat_size.restype = sz
        L.vso_flat_size.argtypes = [vp]
        L.vso_flat_label_count.restype = sz
        L.vso_flat_label_count.argtypes = [vp]
        L.vso_flat_row.restype = vp
        L.vso_flat_row.argtypes = [vp, sz]
        L.vso_flat_label_of.restype = sz
        L.vso_flat_label_of.argtypes = [vp, sz]
        L.vso_flat_topk.restype = sz
        L.vso_flat_topk.argtypes = [vp, vp, sz, i32, i32, vp, vp, C.POINTER(i32)]
        L.vso_flat_range.restype = C.c_long
        L.vso_flat_range.argtypes = [vp, vp, dbl, i32, i32, sz, vp, vp, C.POINTER(i32)]
        L.vso_flat_distance_from.restype = dbl
        L.vso_flat_distance_from.argtypes = [vp, sz, vp]
        L.vso_bi_new.restype = vp
        L.vso_bi_new.argtypes = [vp, vp]
        L.vso_bi_next.restype = sz
        L.vso_bi_next.argtypes = [vp, sz, i32, vp, vp, C.POINTER(i32)]
        L.vso_bi_has_next.argtypes = [vp]
        L.vso_bi_reset.argtypes = [vp]
        L.vso_bi_free.argtypes = [vp]
        L.vso_hnsw_new.restype = vp
        L.vso_hnsw_new.argtypes = [i32, sz, i32, sz, sz, sz, dbl]
        L.vso_hnsw_free.argtypes = [vp]
        L.vso_hnsw_add.argtypes = [vp, vp, sz]
        L.vso_hnsw_size.restype = sz
        L.vso_hnsw_size.argtypes = [vp]
        L.vso_hnsw_set_multi.argtypes = [vp, C.c_int]
        L.vso_hnsw_mark_deleted.argtypes = [vp, sz, i32]
        L.vso_hnsw_info.argtypes = [vp, C.POINTER(C.c_long), C.POINTER(C.c_long)]
        L.vso_hnsw_level.restype = C.c_uint32
        L.vso_hnsw_level.argtypes = [vp, sz]
        L.vso_hnsw_links.restype = sz
        L.vso_hnsw_links.argtypes = [vp, sz, sz, vp]
        L.vso_hnsw_dist_count.restype = sz
        L.vso_hnsw_dist_count.argtypes = [vp]
        L.vso_hnsw_topk.restype = sz
        L.vso_hnsw_topk.argtypes = [vp, vp, sz, sz, vp, vp]
        L.vso_hnsw_range.restype = sz
        L.vso_hnsw_range.argtypes = [vp, vp, dbl, dbl, sz, vp, vp]
        L.vso_hnsw_bi_new.restype = vp
        L.vso_hnsw_bi_new.argtypes = [vp, vp, sz]
        L.vso_hnsw_bi_free.argtypes = [vp]
        L.vso_hnsw_bi_reset.argtypes = [vp]
        L.vso_hnsw_bi_has_next.argtypes = [vp]
        L.vso_hnsw_bi_next.restype = sz
        L.vso_hnsw_bi_next.argtypes = [vp, sz, sz, vp, vp]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def set_tier(tier):
    lib().vso_set_tier(tier)


def stored_size(vtype, metric, dim):
    return lib().vso_stored_size(vtype, metric, dim)


def distance(vtype, metric, a, b, dim=None):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    if dim is None:
        dim = a.size
    return lib().vso_distance(vtype, metric, dim, _ptr(a), _ptr(b))


def distance_many(vtype, metric, dim, A, B):
    A = np.ascontiguousarray(A)
    B = np.ascontiguousarray(B)
    L = lib()
    out = np.empty(A.shape[0], dtype=np.float64)
    for i in range(A.shape[0]):
        out[i] = L.vso_distance(vtype, metric, dim, A[i].ctypes.data, B[i].ctypes.data)
    return out


def normalize(vtype, dim, blob):
    lib().vso_normalize(vtype, dim, _ptr(blob))
    return blob


def to_bf16(x):
    """float32 array -> uint16 bf16 bits, round-to-nearest-even (types/bfloat16.h:23-30)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    u = u + (((u >> 16) & 1) + 0x7FFF)
    return (u >> 16).astype(np.uint16)


def from_bf16(h):
    return (np.ascontiguousarray(h, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


class PortIndex:
    def __init__(self, vtype, dim, metric, multi=False, block_size=1024):
        self.vtype, self.dim, self.metric = vtype, dim, metric
        self.h = lib().vso_flat_new(vtype, dim, metric, int(multi), block_size)

    def close(self):
        if self.h:
            lib().vso_flat_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add(self, blob, label):
        blob = np.ascontiguousarray(blob)
        return lib().vso_flat_add(self.h, _ptr(blob), label)

    def add_many(self, blobs, labels=None, first_label=0):
        blobs = np.ascontiguousarray(blobs)
        L = lib()
        n = 0
        for i in range(blobs.shape[0]):
            n += L.vso_flat_add(self.h, blobs[i].ctypes.data,
                                int(labels[i]) if labels is not None else first_label + i)
        return n

    def delete(self, label):
        return lib().vso_flat_delete(self.h, label)

    def size(self):
        return lib().vso_flat_size(self.h)

    def distance_from(self, label, blob):
        blob = np.ascontiguousarray(blob)
        return lib().vso_flat_distance_from(self.h, label, _ptr(blob))

    def stored_rows(self):
        """(n x stored_bytes) uint8 copy of the processed rows + labels, in internal-id order."""
        n = self.size()
        sb = stored_size(self.vtype, self.metric, self.dim)
        rows = np.empty((n, sb), dtype=np.uint8)
        labels = np.empty(n, dtype=np.uint64)
        L = lib()
        for i in range(n):
            C.memmove(rows[i].ctypes.data, L.vso_flat_row(self.h, i), sb)
            labels[i] = L.vso_flat_label_of(self.h, i)
        return rows, labels

    def topk(self, q, k, order=BY_SCORE, timeout=0):
        q = np.ascontiguousarray(q)
        labels = np.empty(max(k, 1), dtype=np.uint64)
        scores = np.empty(max(k, 1), dtype=np.float64)
        code = C.c_int()
        n = lib().vso_flat_topk(self.h, _ptr(q), k, order, timeout, _ptr(labels), _ptr(scores),
                                C.byref(code))
        return labels[:n].copy(), scores[:n].copy(), code.value

    def range(self, q, radius, order=BY_SCORE, timeout=0):
        q = np.ascontiguousarray(q)
        cap = max(self.size(), 1)
        labels = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        code = C.c_int()
        n = lib().vso_flat_range(self.h, _ptr(q), float(radius), order, timeout, cap, _ptr(labels),
                                 _ptr(scores), C.byref(code))
        if n < 0:
            raise RuntimeError("rangeQuery: invalid radius / order")
        return labels[:n].copy(), scores[:n].copy(), code.value

    def batch_iterator(self, q):
        return PortBatchIterator(self, q)


class PortBatchIterator:
    def __init__(self, index, q):
        q = np.ascontiguousarray(q)
        self.index = index
        self.it = lib().vso_bi_new(index.h, _ptr(q))

    def next(self, n, order=BY_SCORE):
        labels = np.empty(max(n, 1), dtype=np.uint64)
        scores = np.empty(max(n, 1), dtype=np.float64)
        code = C.c_int()
        m = lib().vso_bi_next(self.it, n, order, _ptr(labels), _ptr(scores), C.byref(code))
        return labels[:m].copy(), scores[:m].copy(), code.value

    def has_next(self):
        return bool(lib().vso_bi_has_next(self.it))

    def reset(self):
        lib().vso_bi_reset(self.it)

    def close(self):
        if self.it:
            lib().vso_bi_free(self.it)
            self.it = None


class PortHnsw:
    """vs_oracle_hnsw.c: the reference's single-threaded HNSW build / top-k / range, restated."""

    def __init__(self, vtype, dim, metric, M=16, ef_construction=200, ef_runtime=10, epsilon=0.01, multi=False):
        self.vtype, self.dim, self.metric, self.M = vtype, dim, metric, M
        self.h = lib().vso_hnsw_new(vtype, dim, metric, M, ef_construction, ef_runtime, epsilon)
        self.multi = bool(multi)
        self._label_set = set()
        if multi:  # HNSWIndex_Multi: labels may repeat, queries return each label once
            lib().vso_hnsw_set_multi(self.h, 1)

    def close(self):
        if self.h:
            lib().vso_hnsw_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_many(self, blobs, labels=None, first_label=0):
        blobs = np.ascontiguousarray(blobs)
        L = lib()
        for i in range(blobs.shape[0]):
            label = int(labels[i]) if labels is not None else first_label + i
            L.vso_hnsw_add(self.h, blobs[i].ctypes.data, label)
            if self.multi:
                self._label_set.add(label)
        return blobs.shape[0]

    def size(self):
        return lib().vso_hnsw_size(self.h)

    def label_count(self):
        """indexLabelCount: vectors for a single-value index, distinct labels for a multi-value one."""
        return len(self._label_set) if self.multi else self.size()

    def mark_deleted(self, internal_id, deleted=True):
        lib().vso_hnsw_mark_deleted(self.h, internal_id, int(deleted))

    def topk(self, q, k, ef_runtime=0):
        q = np.ascontiguousarray(q)
        labels = np.empty(max(k, 1), dtype=np.uint64)
        scores = np.empty(max(k, 1), dtype=np.float64)
        n = lib().vso_hnsw_topk(self.h, _ptr(q), k, ef_runtime, _ptr(labels), _ptr(scores))
        return labels[:n].copy(), scores[:n].copy(), 0

    def range(self, q, radius, epsilon=0.0):
        q = np.ascontiguousarray(q)
        cap = max(self.size(), 1)
        labels = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        n = lib().vso_hnsw_range(self.h, _ptr(q), float(radius), float(epsilon), cap, _ptr(labels), _ptr(scores))
        return labels[:n].copy(), scores[:n].copy(), 0

    def export(self):
        """Same layout as oracle.ref.RefIndex.hnsw_export (levels, links[l] [n, width], counts[l] [n])."""
        L = lib()
        n, M = self.size(), self.M
        entry, maxl = C.c_long(), C.c_long()
        L.vso_hnsw_info(self.h, C.byref(entry), C.byref(maxl))
        levels = np.array([L.vso_hnsw_level(self.h, i) for i in range(n)], dtype=np.uint32)
        out = dict(n=n, M=M, entry=entry.value, max_level=maxl.value, levels=levels, links=[], counts=[])
        buf = np.empty(2 * M, dtype=np.uint32)
        for lvl in range(max(maxl.value, 0) + 1):
            width = 2 * M if lvl == 0 else M
            links = np.full((n, width), 0xFFFFFFFF, dtype=np.uint32)
            counts = np.zeros(n, dtype=np.uint32)
            for i in range(n):
                if levels[i] < lvl:
                    continue
                c = L.vso_hnsw_links(self.h, i, lvl, _ptr(buf))
                counts[i] = c
                links[i, :c] = buf[:c]
            out["links"].append(links)
            out["counts"].append(counts)
        return out

    def dist_count(self):
        return lib().vso_hnsw_dist_count(self.h)

    def batch_iterator(self, q, ef_runtime=0):
        return PortHnswBatchIterator(self, q, ef_runtime)


class PortHnswBatchIterator:
    def __init__(self, index, q, ef_runtime=0):
        q = np.ascontiguousarray(q)
        self.index = index
        self.it = lib().vso_hnsw_bi_new(index.h, _ptr(q), ef_runtime)

    def next(self, n, order=BY_SCORE):
        labels = np.empty(max(n, 1), dtype=np.uint64)
        scores = np.empty(max(n, 1), dtype=np.float64)
        m = lib().vso_hnsw_bi_next(self.it, n, self.index.label_count(), _ptr(labels), _ptr(scores))
        labels, scores = labels[:m].copy(), scores[:m].copy()
        if order == BY_ID:
            o = np.argsort(labels, kind="stable")
            labels, scores = labels[o], scores[o]
        return labels, scores, 0

    def has_next(self):
        return bool(lib().vso_hnsw_bi_has_next(self.it))

    def reset(self):
        lib().vso_hnsw_bi_reset(self.it)

    def close(self):
        if self.it:
            lib().vso_hnsw_bi_free(self.it)
            self.it = None


# ---- tiered index: merge of the two tiers' replies -------------------------------------------------------------------
VECSIM_EPSILON = 1e-6  # /root/reference/src/VecSim/utils/query_result_utils.h:14


def _cmp_score_then_id(a, b):
    """cmpVecSimQueryResultByScoreThenId (query_result_utils.h:18-23); a, b = (id, score)."""
    if not abs(a[1] - b[1]) < VECSIM_EPSILON:
        return 1 if a[1] > b[1] else -1
    d = (int(a[0]) - int(b[0])) & 0xFFFFFFFFFFFFFFFF  # size_t difference ...
    d &= 0xFFFFFFFF                                     # ... truncated to int
    return d - (1 << 32) if d & 0x80000000 else d


def merge_results(first, second, limit):
    """merge_results<withSet=false> (query_result_utils.h:44-92): both lists ascending by (score, id), entries are
    (id, score); a result present in both (same score) is emitted once. -> (merged, taken_first, taken_second)."""
    out, i, j = [], 0, 0
    limit = (1 << 64) - 1 if limit is None or limit < 0 else limit
    while limit and i < len(first) and j < len(second):
        c = _cmp_score_then_id(first[i], second[j])
        if c > 0:
            out.append(second[j])
            j += 1
        elif c < 0:
            out.append(first[i])
            i += 1
        else:
            out.append(first[i])
            i += 1
            j += 1
        limit -= 1
    if limit:
        if i == len(first):
            while limit and j < len(second):
                out.append(second[j])
                j += 1
                limit -= 1
        else:
            while limit and i < len(first):
                out.append(first[i])
                i += 1
                limit -= 1
    return out, i, j


# ---- HNSW index files (encoding V3 / V4) -----------------------------------------------------------------------------
def read_hnsw_file(path):
    """Parse a serialized HNSW index the way the reference restores it
    (/root/reference/src/VecSim/index_factories/hnsw_factory.cpp:171-205 header,
    algorithms/hnsw/hnsw_serializer_impl.h:145-245 fields / metadata / graph,
    containers/data_blocks_container.cpp:76-110 vector blocks; V3 stores block counts and lengths, V4 does not).
    -> dict(version, params..., labels[n], flags[n], vectors[n, stored bytes], levels[n], links[l][n, width],
    counts[l][n], incoming (number of unidirectional incoming edges, unused by searches))."""
    import struct
    buf = open(path, "rb").read()
    pos = 0

    def rd(fmt):
        nonlocal pos
        v = struct.unpack_from("<" + fmt, buf, pos)
        pos += struct.calcsize("<" + fmt)
        return v[0] if len(v) == 1 else v
    version = rd("i")
    if version not in (3, 4):
        raise ValueError("encoding version %d is not V3 / V4" % version)
    algo = rd("i")
    if algo != 1:
        raise ValueError("not an HNSW file (algo %d)" % algo)
    dim, vtype, metric, block_size = rd("Q"), rd("i"), rd("i"), rd("Q")
    multi, cap = rd("?"), rd("Q")
    M, M0, efc, ef = rd("Q"), rd("Q"), rd("Q"), rd("Q")
    epsilon, mult = rd("d"), rd("d")
    n, n_deleted, max_level, entry = rd("Q"), rd("Q"), rd("Q"), rd("I")
    labels = np.empty(n, dtype=np.uint64)
    flags = np.empty(n, dtype=np.uint8)
    for i in range(n):
        labels[i], flags[i] = rd("Q"), rd("B")
    stored = stored_size(vtype, metric, dim)
    vectors = np.empty((n, stored), dtype=np.uint8)
    if version == 3:
        num_blocks = rd("I")
        got = 0
        for _ in range(num_blocks):
            blen = rd("I")
            vectors[got:got + blen] = np.frombuffer(buf, np.uint8, blen * stored, pos).reshape(blen, stored)
            pos += blen * stored
            got += blen
        assert got == n
    else:
        num_blocks = -(-n // block_size) if block_size else 0
        vectors[:] = np.frombuffer(buf, np.uint8, n * stored, pos).reshape(n, stored)
        pos += n * stored
    levels = np.zeros(n, dtype=np.uint32)
    per_level = {}
    incoming = 0
    i = 0
    for _ in range(num_blocks):
        blen = rd("I")
        for _ in range(blen):
            top = rd("Q")
            levels[i] = top
            for lvl in range(top + 1):
                cnt = rd("H")
                ids = struct.unpack_from("<%dI" % cnt, buf, pos)
                pos += 4 * cnt
                inc = rd("I")
                pos += 4 * inc
                incoming += inc
                per_level.setdefault(lvl, {})[i] = ids
            i += 1
    assert i == n and pos == len(buf), (i, n, pos, len(buf))
    nl = (max(per_level) + 1) if per_level else 0
    links, counts = [], []
    for lvl in range(nl):
        width = M0 if lvl == 0 else M
        lk = np.full((n, width), 0xFFFFFFFF, dtype=np.uint32)
        ct = np.zeros(n, dtype=np.uint32)
        for j, ids in per_level[lvl].items():
            ct[j] = len(ids)
            lk[j, :len(ids)] = ids
        links.append(lk)
        counts.append(ct)
    return dict(version=version, dim=dim, type=vtype, metric=metric, block_size=block_size, multi=multi, capacity=cap, M=M, M0=M0,
                ef_construction=efc, ef_runtime=ef, epsilon=epsilon, mult=mult, n=n, num_deleted=n_deleted,
                max_level=None if max_level == 0xFFFFFFFFFFFFFFFF else max_level,
                entry=None if entry == 0xFFFFFFFF else entry, labels=labels, flags=flags, vectors=vectors, levels=levels,
                links=links, counts=counts, incoming=incoming)
