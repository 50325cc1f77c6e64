"""Multi-value HNSW on the device (HNSWIndex_Multi, SURVEY §8 row f2, second half) through the C API: several vectors per
label, queries return each label once with its best score. Compared bit for bit — labels, order, scores — with the oracle
port, whose multi-value behaviour tests/test_oracle_vs_reference.py pins to the unmodified reference
(test_hnsw_multi_port_matches_reference, ..._batch_iterator_matches_reference). The device builder inserts in the same
order as the oracle, so both search the same graph. Mirrors tests/unit/test_hnsw_multi.cpp."""
import numpy as np
import pytest

from datagen import make_vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    return c


def new_multi(capi, vtype, dim, metric, M=8, efc=40, ef=20):
    return capi.HNSWIndex(capi.HNSWParams(type=vtype, dim=dim, metric=metric, multi=True, initialCapacity=0, blockSize=1024,
                                          M=M, efConstruction=efc, efRuntime=ef, epsilon=0.01))


@pytest.mark.parametrize("vtype,metric", [(0, 0), (0, 1), (0, 2), (2, 0), (3, 1), (4, 2), (5, 0), (1, 0)], ids=lambda v: str(v))
def test_multi_topk_and_range_match_oracle(capi, port, vtype, metric):
    port.set_tier(port.TIER_AVX512)
    n, dim = 1200, 16
    labels = (np.arange(n) % 150).astype(np.uint64)              # eight vectors per label
    X = make_vectors(vtype, n, dim, seed=3)
    Q = make_vectors(vtype, 10, dim, seed=4)
    if metric == 2:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    G = new_multi(capi, vtype, dim, metric)
    assert G.add_vectors(X, labels=labels) == n                  # a repeated label is a NEW vector, never an overwrite
    assert G.index_size() == n
    info = dict(G.debug_info())
    assert info["IS_MULTI_VALUE"] == 1 and info["INDEX_LABEL_COUNT"] == 150
    P = port.PortHnsw(vtype, dim, metric, M=8, ef_construction=40, ef_runtime=20, multi=True)
    P.add_many(X, labels=labels)
    for k, ef in ((10, 20), (5, 50), (40, 20), (150, 20), (30, 100)):     # ef <= 64: register-resident set; 100: shared memory
        G.set_ef(ef)
        gl, gs = G.knn_batch(Q, k)
        for i, q in enumerate(Q):
            pl, ps, _ = P.topk(q, k, ef_runtime=ef)
            got_l, got_s = gl[i][:len(pl)], gs[i][:len(pl)]
            assert len(set(got_l.tolist())) == len(got_l)                 # each label once
            assert np.array_equal(got_l, pl.astype(np.int64)) and np.array_equal(got_s, ps), (vtype, metric, k, ef, i)
            assert (gl[i][len(pl):] == -1).all()
    G.set_ef(20)
    for q in Q[:5]:
        s20 = P.topk(q, 20)[1]
        radius = float(s20[-1]) if s20[-1] > 0 else 0.3
        pl, ps, _ = P.range(q, radius)
        gl, gs = G.range_query(q, radius)
        assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps), (vtype, metric, "range")
    # distance from a label = its closest vector; deleting a label removes all of its vectors
    if metric != 2:   # (cosine stores normalised rows and takes the caller's blob as it is: nothing to compare with here)
        want = min(port.distance(vtype, metric, X[i], Q[0]) for i in range(7, n, 150))
        assert G.get_distance_from(7, Q[0]) == want
    assert G.delete_vector(7) == 8
    assert G.index_size() == n - 8 and dict(G.debug_info())["INDEX_LABEL_COUNT"] == 149
    gl, _ = G.knn_batch(Q, 149)
    assert not (gl == 7).any()
    G.close()
    P.close()


@pytest.mark.parametrize("metric", [0, 2], ids=["L2", "Cosine"])
def test_multi_batch_iterator_matches_oracle(capi, port, metric):
    port.set_tier(port.TIER_AVX512)
    n, dim = 900, 16
    labels = (np.arange(n) % 100).astype(np.uint64)               # nine vectors per label
    X = make_vectors(0, n, dim, seed=81 + metric)
    Q = make_vectors(0, 4, dim, seed=82 + metric)
    G = new_multi(capi, 0, dim, metric, M=6, efc=40, ef=10)
    G.add_vectors(X, labels=labels)
    P = port.PortHnsw(0, dim, metric, M=6, ef_construction=40, ef_runtime=10, multi=True)
    P.add_many(X, labels=labels)
    for q in Q:
        for sched in ([5, 5, 5, 20, 1, 100], [1, 2, 3], [50, 50], [1000]):
            gi, pi = G.create_batch_iterator(q), P.batch_iterator(q)
            for rounds in range(2):
                seen = []
                for nres in sched:
                    assert gi.has_next() == pi.has_next()
                    gl, gs = gi.get_next_results(nres)
                    pl, ps, _ = pi.next(nres)
                    assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps), (sched, nres)
                    seen += pl.tolist()
                assert len(seen) == len(set(seen))
                assert gi.has_next() == pi.has_next()
                gi.reset()
                pi.reset()
            gi.close()
            pi.close()
    G.close()
    P.close()


def test_multi_index_file_round_trip(capi, port, tmp_path):
    """A multi-value index saved as V4 and loaded again answers identically (labels with several vectors, one label
    deleted before the save)."""
    n, dim = 600, 8
    labels = (np.arange(n) % 75).astype(np.uint64)
    X = make_vectors(0, n, dim, seed=9)
    Q = make_vectors(0, 6, dim, seed=10)
    G = new_multi(capi, 0, dim, 0)
    G.add_vectors(X, labels=labels)
    assert G.delete_vector(11) == 8
    path = str(tmp_path / "multi.hnsw_v4")
    G.save_index(path)
    G2 = capi.HNSWIndex.load(path)
    assert dict(G2.debug_info())["IS_MULTI_VALUE"] == 1
    assert G2.index_size() == G.index_size() == n - 8
    a_l, a_s = G.knn_batch(Q, 30)
    b_l, b_s = G2.knn_batch(Q, 30)
    assert np.array_equal(a_l, b_l) and np.array_equal(a_s, b_s)
    f = port.read_hnsw_file(path)
    assert f["multi"] and f["n"] == n and f["num_deleted"] == 8 and int((f["flags"] & 1).sum()) == 8
    G.close()
    G2.close()


def test_tiered_index_over_multi_value_backend(capi, port):
    """VecSimAlgo_TIERED over a multi-value HNSW backend (hnsw_tiered.h with HNSWIndex_Multi / BruteForceIndex_Multi): a
    repeated label is another vector; while some vectors of a label are still buffered and others already ingested, a query
    returns the label once, with the best score over both tiers (merge_results<withSet = true>,
    utils/query_result_utils.h:44-92). Checked against the oracle halves + the restated merge, and — once everything is
    ingested — against the oracle's multi-value HNSW built in the same order."""
    n, dim, n_labels = 600, 16, 60
    labels = (np.arange(n) % n_labels).astype(np.uint64)          # ten vectors per label
    X = make_vectors(0, n, dim, seed=21)
    Q = make_vectors(0, 8, dim, seed=22)
    T = capi.Tiered_HNSWIndex(capi.HNSWParams(type=0, dim=dim, metric=0, multi=True, initialCapacity=0, blockSize=1024, M=8,
                                              efConstruction=40, efRuntime=60, epsilon=0.01), None, flat_buffer_size=1000)
    for i in range(n):
        assert T.add_vector(X[i], int(labels[i])) == 1
    assert T.index_size() == n and T.get_curr_bf_size() == n
    # everything buffered: the flat tier answers alone (exact multi-value scan)
    PF = port.PortIndex(0, dim, 0, multi=True)
    PF.add_many(X, labels=labels)
    for q in Q:
        l, s = T.knn_query(q, 10)
        pl, ps, _ = PF.topk(q, 10)
        assert np.array_equal(l[0], pl.astype(np.int64)) and np.array_equal(s[0], ps)
    # one executed job drains everything pending at that moment into the backend, in submission order
    T.run_jobs(max_jobs=1)
    assert T.get_curr_bf_size() == 0 and T.index_size() == n
    PB = port.PortHnsw(0, dim, 0, M=8, ef_construction=40, ef_runtime=60, multi=True)
    PB.add_many(X, labels=labels)
    for q in Q:
        l, s = T.knn_query(q, 10)
        pl, ps, _ = PB.topk(q, 10, ef_runtime=60)
        assert np.array_equal(l[0], pl.astype(np.int64)) and np.array_equal(s[0], ps)
        assert len(set(l[0].tolist())) == 10
    # a second wave stays in the buffer: labels now live in BOTH tiers
    X2 = make_vectors(0, 120, dim, seed=23)
    lab2 = (np.arange(120) % n_labels).astype(np.uint64)
    for i in range(120):
        assert T.add_vector(X2[i], int(lab2[i])) == 1
    assert T.get_curr_bf_size() == 120 and T.index_size() == n + 120
    PF2 = port.PortIndex(0, dim, 0, multi=True)
    PF2.add_many(X2, labels=lab2)
    for q in Q:
        fl, fs, _ = PF2.topk(q, 10)
        bl, bs, _ = PB.topk(q, 10, ef_runtime=60)
        want = {}
        for lab, sc in list(zip(fl.tolist(), fs.tolist())) + list(zip(bl.tolist(), bs.tolist())):
            want[lab] = min(sc, want.get(lab, np.inf))
        want = sorted(want.items(), key=lambda t: (t[1], t[0]))[:10]
        l, s = T.knn_query(q, 10)
        assert l[0].tolist() == [w[0] for w in want] and s[0].tolist() == [w[1] for w in want]
        assert len(set(l[0].tolist())) == 10
    # deleting a label removes its vectors from both tiers and voids its pending jobs
    removed = T.delete_vector(7)
    assert removed == 10 + 2 and T.index_size() == n + 120 - 12
    l, _ = T.knn_query(X[7], 60)
    assert 7 not in l[0].tolist()
    T.wait_for_index()
    assert T.get_curr_bf_size() == 0
    T.close()
    PF.close()
    PF2.close()
    PB.close()
