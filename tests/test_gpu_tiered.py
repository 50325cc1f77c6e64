"""Tiered index (flat buffer in front of an HNSW backend) on the device, through the C-ABI.
Expected results are composed from the oracle's two halves — the exact flat scan over what sits in the buffer and the
HNSW build + traversal over what was ingested, in ingestion order — merged with the restated merge_results
(query_result_utils.h:44-92), the way vec_sim_tiered_index.h:169-316 composes them. Labels, order and fp scores are
compared bit-exact. Mirrors tests/flow/test_hnsw_tiered.py and tests/unit/test_hnsw_tiered.cpp of the reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
DIM, M, EFC, EF, K = 32, 8, 40, 20, 10


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    return c


@pytest.fixture(scope="module")
def port():
    from oracle import port as p
    p.build()
    return p


def hnsw_params(capi, vtype=0, metric=0):
    return capi.HNSWParams(type=vtype, dim=DIM, metric=metric, multi=False, initialCapacity=0, blockSize=1024, M=M,
                           efConstruction=EFC, efRuntime=EF, epsilon=0.01)


def new_tiered(capi, limit, vtype=0, metric=0, threads=0):
    return capi.Tiered_HNSWIndex(hnsw_params(capi, vtype, metric), None, flat_buffer_size=limit, threads=threads)


def expected_topk(port, flat, hnsw, q, k):
    fl, fs, _ = flat.topk(q, k) if flat is not None and flat.size() else (np.zeros(0, np.int64), np.zeros(0), 0)
    hl, hs, _ = hnsw.topk(q, k) if hnsw is not None and hnsw.size() else (np.zeros(0, np.int64), np.zeros(0), 0)
    a = sorted(zip(hl.tolist(), hs.tolist()), key=lambda t: (t[1], t[0]))
    b = sorted(zip(fl.tolist(), fs.tolist()), key=lambda t: (t[1], t[0]))
    merged, _, _ = port.merge_results(a, b, k)
    return np.array([m[0] for m in merged], dtype=np.int64), np.array([m[1] for m in merged])


@pytest.mark.parametrize("metric", [0, 2])
def test_buffered_then_ingested(capi, port, metric):
    """Everything sits in the flat buffer until the jobs run (answers = the exact scan), afterwards everything is in
    the backend (answers = HNSW built in submission order)."""
    rng = np.random.default_rng(1)
    n = 600
    X = rng.uniform(-1, 1, (n, DIM)).astype(np.float32)
    Q = rng.uniform(-1, 1, (6, DIM)).astype(np.float32)
    T = new_tiered(capi, limit=1000, metric=metric)
    for i in range(n):
        assert T.add_vector(X[i], i + 7) == 1
    assert T.index_size() == n and T.get_curr_bf_size() == n and T.pending_jobs() == n
    F = port.PortIndex(0, DIM, metric)
    F.add_many(X, first_label=7)
    for q in Q:
        wl, ws = expected_topk(port, F, None, q, K)
        l, s = T.knn_query(q, K)
        assert np.array_equal(l[0], wl) and np.array_equal(s[0], ws)
    info = dict(T.debug_info())
    assert info["ALGORITHM"] == "TIERED" and info["BACKGROUND_INDEXING"] == 1 and info["TIERED_BUFFER_LIMIT"] == 1000
    assert dict(info["FRONTEND_INDEX"])["ALGORITHM"] == "FLAT" and dict(info["FRONTEND_INDEX"])["INDEX_SIZE"] == n
    assert dict(info["BACKEND_INDEX"])["ALGORITHM"] == "HNSW" and dict(info["BACKEND_INDEX"])["INDEX_SIZE"] == 0
    T.wait_for_index()
    assert T.pending_jobs() == 0 and T.get_curr_bf_size() == 0 and T.index_size() == n and T.hnsw_label_count() == n
    H = port.PortHnsw(0, DIM, metric, M=M, ef_construction=EFC, ef_runtime=EF)
    H.add_many(X, first_label=7)
    for q in Q:
        wl, ws = expected_topk(port, None, H, q, K)
        l, s = T.knn_query(q, K)
        assert np.array_equal(l[0], wl) and np.array_equal(s[0], ws)
    info = dict(T.debug_info())
    assert info["BACKGROUND_INDEXING"] == 0 and info["INDEX_SIZE"] == n and info["INDEX_LABEL_COUNT"] == n
    assert T.stats()["directHNSWInsertions"] == 0
    F.close(); H.close(); T.close()


@pytest.mark.parametrize("vtype", [0, 2, 4])
def test_both_tiers_merge(capi, port, vtype):
    """Buffer limit 150: the first 150 vectors wait in the flat buffer, the rest go straight to the backend; top-K,
    batched top-K and range queries merge both tiers."""
    from datagen import make_vectors
    n, limit = 500, 150
    metric = 2 if vtype == 4 else 0
    X = make_vectors(vtype, n, DIM, seed=21)
    Q = make_vectors(vtype, 5, DIM, seed=22)
    T = new_tiered(capi, limit=limit, vtype=vtype, metric=metric)
    for i in range(n):
        assert T.add_vector(X[i], i) == 1
    st = T.stats()
    assert st["flatBufferSize"] == limit and st["directHNSWInsertions"] == n - limit and T.index_size() == n
    F = port.PortIndex(vtype, DIM, metric)
    F.add_many(X[:limit], first_label=0)
    H = port.PortHnsw(vtype, DIM, metric, M=M, ef_construction=EFC, ef_runtime=EF)
    H.add_many(X[limit:], first_label=limit)
    want = [expected_topk(port, F, H, q, K) for q in Q]
    for q, (wl, ws) in zip(Q, want):
        l, s = T.knn_query(q, K)
        assert np.array_equal(l[0], wl) and np.array_equal(s[0], ws)
    bl, bs = T.knn_batch(Q, K)
    for i, (wl, ws) in enumerate(want):
        assert np.array_equal(bl[i][:len(wl)], wl) and np.array_equal(bs[i][:len(ws)], ws)
    # range: union of both tiers, by score and by id
    q = Q[0]
    _, fs, _ = F.topk(q, 40)
    radius = float(fs[-1])
    fl, fsc, _ = F.range(q, radius)
    hl, hsc, _ = H.range(q, radius)
    a = sorted(zip(hl.tolist(), hsc.tolist()), key=lambda t: (t[1], t[0]))
    b = sorted(zip(fl.tolist(), fsc.tolist()), key=lambda t: (t[1], t[0]))
    merged, _, _ = port.merge_results(a, b, None)
    l, s = T.range_query(q, radius)
    assert l[0].tolist() == [m[0] for m in merged] and s[0].tolist() == [m[1] for m in merged]
    l, s = T.range_query(q, radius, order=capi.BY_ID)
    by_id = sorted(merged)
    assert l[0].tolist() == [m[0] for m in by_id] and s[0].tolist() == [m[1] for m in by_id]
    # distances to labels in either tier (the ad-hoc context preprocesses the query like a search does)
    el, es, _ = F.topk(q, limit)
    exact = dict(zip(el.tolist(), es.tolist()))
    d = T.adhoc_distances(q, [3, limit + 3, 10 ** 6])
    assert d[0] == exact[3] and np.isnan(d[2]) and not np.isnan(d[1])
    if vtype != 4:  # int8 cosine: GetDistanceFrom_Unsafe wants the dim+4-byte processed blob, covered by the flat tests
        assert T.get_distance_from(3, q) == F.distance_from(3, q)
        assert not np.isnan(T.get_distance_from(limit + 3, q)) and np.isnan(T.get_distance_from(10 ** 6, q))
    F.close(); H.close(); T.close()


def test_overwrite_and_delete(capi, port):
    rng = np.random.default_rng(3)
    X = rng.uniform(-1, 1, (200, DIM)).astype(np.float32)
    T = new_tiered(capi, limit=1000)
    for i in range(100):
        assert T.add_vector(X[i], i) == 1
    # overwrite in the buffer: the old job is void, the label keeps one vector
    assert T.add_vector(X[150], 5) == 0
    assert T.index_size() == 100 and T.get_curr_bf_size() == 100
    # delete from the buffer
    assert T.delete_vector(6) == 1 and T.delete_vector(6) == 0 and T.delete_vector(10 ** 6) == 0
    assert T.index_size() == 99
    T.wait_for_index()
    assert T.get_curr_bf_size() == 0 and T.index_size() == 99
    l, s = T.knn_query(X[150], 1)
    assert l[0][0] == 5 and s[0][0] == 0.0
    l, _ = T.knn_query(X[6], 99)
    assert 6 not in l[0].tolist() and len(l[0]) >= 90
    # overwrite after ingestion: new vector buffered, old one tombstoned in the backend at once
    assert T.add_vector(X[151], 7) == 0
    assert T.get_curr_bf_size() == 1 and T.index_size() == 99
    assert T.stats()["numberOfMarkedDeleted"] >= 1
    l, s = T.knn_query(X[151], 3)
    assert l[0][0] == 7 and s[0][0] == 0.0 and l[0].tolist().count(7) == 1
    l, _ = T.knn_query(X[7], 99)
    assert l[0].tolist().count(7) == 1
    # delete a label that lives in the backend
    assert T.delete_vector(20) == 1
    assert T.index_size() == 98
    T.wait_for_index()
    assert T.get_curr_bf_size() == 0 and T.index_size() == 98
    l, _ = T.knn_query(X[20], 98)
    got = l[0].tolist()
    assert 20 not in got and 6 not in got and len(set(got)) == len(got) and len(got) >= 90
    T.close()


def test_write_in_place(capi, port):
    rng = np.random.default_rng(4)
    X = rng.uniform(-1, 1, (120, DIM)).astype(np.float32)
    T = new_tiered(capi, limit=1000)
    capi.set_write_mode(True)
    try:
        for i in range(120):
            assert T.add_vector(X[i], i) == 1
        assert T.add_vector(X[0], 5) == 0
    finally:
        capi.set_write_mode(False)
    assert T.get_curr_bf_size() == 0 and T.pending_jobs() == 0 and T.index_size() == 120
    assert T.stats()["directHNSWInsertions"] == 121
    l, s = T.knn_query(X[0], 2)
    assert sorted(l[0].tolist()) == [0, 5] and s[0][0] == 0.0 and s[0][1] == 0.0
    T.close()


def test_batch_iterator_covers_both_tiers_once(capi, port):
    rng = np.random.default_rng(5)
    n, limit = 400, 120
    X = rng.uniform(-1, 1, (n, DIM)).astype(np.float32)
    q = rng.uniform(-1, 1, DIM).astype(np.float32)
    T = new_tiered(capi, limit=limit)
    for i in range(n):
        T.add_vector(X[i], i)
    F = port.PortIndex(0, DIM, 0)
    F.add_many(X, first_label=0)
    exact_l, exact_s, _ = F.topk(q, n)
    score_of = dict(zip(exact_l.tolist(), exact_s.tolist()))
    it = T.create_batch_iterator(q)
    seen, last = [], -1.0
    rounds = 0
    while it.has_next() and rounds < 200:
        l, s = it.get_next_results(25)
        rounds += 1
        assert len(l[0]) <= 25
        for lab, sc in zip(l[0].tolist(), s[0].tolist()):
            assert sc == score_of[lab]
        assert np.all(np.diff(s[0]) >= 0)
        seen += l[0].tolist()
    assert len(seen) == len(set(seen)), "a label came back twice"
    # the flat tier is exact, the backend approximate: everything buffered must be there, and nearly all of the rest
    assert set(range(limit)) <= set(seen) and len(seen) >= int(0.95 * n)
    # the first batch is the merge of the two tiers' first batches
    it.reset()
    l, s = it.get_next_results(10)
    H = port.PortHnsw(0, DIM, 0, M=M, ef_construction=EFC, ef_runtime=EF)
    H.add_many(X[limit:], first_label=limit)
    F2 = port.PortIndex(0, DIM, 0)
    F2.add_many(X[:limit], first_label=0)
    hit = H.batch_iterator(q)
    hl, hs, _ = hit.next(10)
    fl, fs, _ = F2.topk(q, 10)
    a = sorted(zip(hl.tolist(), hs.tolist()), key=lambda t: (t[1], t[0]))
    b = sorted(zip(fl.tolist(), fs.tolist()), key=lambda t: (t[1], t[0]))
    merged, _, _ = port.merge_results(a, b, 10)
    assert l[0].tolist() == [m[0] for m in merged] and s[0].tolist() == [m[1] for m in merged]
    it.close(); hit.close(); H.close(); F.close(); F2.close(); T.close()


def test_worker_threads_drain_while_queries_run(capi, port):
    """Jobs run on four threads while the main thread keeps querying: every answer is the exact nearest neighbour of a
    stored vector queried by itself, whichever tier holds it at that moment."""
    import threading
    rng = np.random.default_rng(6)
    n = 800
    X = rng.uniform(-1, 1, (n, DIM)).astype(np.float32)
    T = new_tiered(capi, limit=10000, threads=4)
    for i in range(n):
        T.add_vector(X[i], i)
    stop = threading.Event()
    errors = []

    def query_loop():
        j = 0
        while not stop.is_set():
            l, s = T.knn_query(X[j % n], 1)
            if l[0][0] != j % n or s[0][0] != 0.0:
                errors.append((j % n, l[0].tolist(), s[0].tolist()))
            j += 37
    t = threading.Thread(target=query_loop)
    t.start()
    T.wait_for_index(60)
    stop.set()
    t.join()
    assert not errors, errors[:3]
    assert T.get_curr_bf_size() == 0 and T.index_size() == n and T.pending_jobs() == 0
    T.close()


def test_empty_k0_timeout_and_locks(capi, port):
    """Edge cases of tests/unit/test_hnsw_tiered.cpp: an empty tiered index answers with empty replies, k = 0 too; a
    firing timeout callback gives an empty TimedOut reply from whichever tier is asked first; the shared-lock API
    brackets reads."""
    rng = np.random.default_rng(8)
    X = rng.uniform(-1, 1, (300, DIM)).astype(np.float32)
    q = X[3]
    T = new_tiered(capi, limit=100)
    l, s = T.knn_query(q, 5)
    assert l.shape == (1, 0) and T.last_code == capi.VecSim_QueryReply_OK
    l, s = T.range_query(q, 1.0)
    assert l.shape == (1, 0)
    assert T.index_size() == 0 and dict(T.debug_info())["INDEX_LABEL_COUNT"] == 0
    for i in range(300):
        T.add_vector(X[i], i)          # 100 buffered, 200 direct
    l, s = T.knn_query(q, 0)
    assert l.shape == (1, 0)
    L = capi.lib()
    L.VecSimTieredIndex_AcquireSharedLocks(T._h)
    d = T.get_distance_from(3, q)
    L.VecSimTieredIndex_ReleaseSharedLocks(T._h)
    assert d == 0.0
    L.VecSimTieredIndex_GC(T._h)
    capi.set_timeout_callback(lambda ctx: 1)
    try:
        l, s = T.knn_query(q, 5)
        assert l.shape == (1, 0) and T.last_code == capi.VecSim_QueryReply_TimedOut
        T.wait_for_index()
        l, s = T.knn_query(q, 5)       # buffer empty now: the backend alone times out
        assert l.shape == (1, 0) and T.last_code == capi.VecSim_QueryReply_TimedOut
    finally:
        capi.set_timeout_callback(None)
    l, s = T.knn_query(q, 5)
    assert l[0][0] == 3 and s[0][0] == 0.0 and T.last_code == capi.VecSim_QueryReply_OK
    # prefer-ad-hoc follows the bigger tier; a small subset is scored ad hoc
    assert T.prefer_adhoc(3, 5) is True
    T.close()


def test_cosine_tiered_matches_halves(capi, port):
    """Cosine: both tiers normalise the caller's blob themselves, so a vector scores the same wherever it sits."""
    rng = np.random.default_rng(9)
    n, limit = 300, 90
    X = rng.uniform(-1, 1, (n, DIM)).astype(np.float32)
    Q = rng.uniform(-1, 1, (4, DIM)).astype(np.float32)
    T = new_tiered(capi, limit=limit, metric=2)
    for i in range(n):
        T.add_vector(X[i], i)
    F = port.PortIndex(0, DIM, 2)
    F.add_many(X[:limit], first_label=0)
    H = port.PortHnsw(0, DIM, 2, M=M, ef_construction=EFC, ef_runtime=EF)
    H.add_many(X[limit:], first_label=limit)
    for q in Q:
        wl, ws = expected_topk(port, F, H, q, K)
        l, s = T.knn_query(q, K)
        assert np.array_equal(l[0], wl) and np.array_equal(s[0], ws)
    # after ingestion the buffered vectors were appended to the backend in submission order
    T.wait_for_index()
    H.add_many(X[:limit], first_label=0)
    for q in Q:
        wl, ws = expected_topk(port, None, H, q, K)
        l, s = T.knn_query(q, K)
        assert np.array_equal(l[0], wl) and np.array_equal(s[0], ws)
    F.close(); H.close(); T.close()
