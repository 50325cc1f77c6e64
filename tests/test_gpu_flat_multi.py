"""Multi-value flat index (BruteForceIndex_Multi) through the C API, against the oracle port
(tests/test_oracle_vs_reference.py pins the port's multi-value behaviour to the unmodified reference)."""
import numpy as np
import pytest

from datagen import METRIC_NAMES, TYPE_NAMES, make_vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    c.set_topk_mode(0)
    return c


def make_pair(capi, port, vtype, dim, metric, X, labels):
    G = capi.BFIndex(capi.BFParams(type=vtype, dim=dim, metric=metric, multi=True, initialCapacity=0, blockSize=1024))
    P = port.PortIndex(vtype, dim, metric, multi=True)
    assert G.add_vectors(X, labels=labels) == len(X)
    P.add_many(X, labels=labels)
    return G, P


@pytest.mark.parametrize("vtype,metric,dim", [(0, 0, 32), (0, 2, 48), (2, 1, 64), (4, 2, 40), (1, 0, 9), (5, 0, 33)],
                         ids=lambda v: str(v))
def test_multi_topk_range_batches_match_oracle(capi, port, vtype, metric, dim):
    port.set_tier(port.TIER_AVX512)
    n, n_labels, nq = 3000, 700, 12
    rng = np.random.default_rng(11 + vtype)
    X = make_vectors(vtype, n, dim, seed=50 + vtype + metric)
    Q = make_vectors(vtype, nq, dim, seed=51 + vtype + metric)
    if metric == 2:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    labels = rng.integers(0, n_labels, n).astype(np.uint64)          # ~4 vectors per label, some labels unused
    G, P = make_pair(capi, port, vtype, dim, metric, X, labels)
    assert G.index_size() == P.size() == n
    for k in (1, 10, 200, 5000):
        gl, gs = G.knn_batch(Q, k)
        for i in range(nq):
            pl, ps, _ = P.topk(Q[i], k)
            m = len(pl)
            assert np.array_equal(gl[i][:m], pl.astype(np.int64)), (TYPE_NAMES[vtype], METRIC_NAMES[metric], k, i)
            assert np.array_equal(gs[i][:m], ps)
            assert (gl[i][m:] == -1).all()
    for i in range(4):
        radius = max(float(P.topk(Q[i], 40)[1][-1]), 0.0)
        for order in (0, 1):
            gl, gs = G.range_query(Q[i], radius, order=order)
            pl, ps, _ = P.range(Q[i], radius, order=order)
            if order == 0:
                o = np.lexsort((pl, ps))
                pl, ps = pl[o], ps[o]
            assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps)
    # batch iterator: each label once, by its best vector
    it = G.create_batch_iterator(Q[0])
    pit = P.batch_iterator(Q[0])
    got = []
    while it.has_next():
        l, s = it.get_next_results(97)
        pl, ps, _ = pit.next(97)
        assert sorted(zip(s[0], l[0])) == sorted(zip(ps, pl.astype(np.int64)))
        got += list(l[0])
    assert len(got) == len(set(got)) == len(np.unique(labels))
    it.close()
    pit.close()
    # getDistanceFrom: minimum over the label's vectors; delete removes every vector of the label
    lab = int(labels[5])
    assert G.get_distance_from(lab, Q[1] if metric != 2 else X[0]) == P.distance_from(lab, Q[1] if metric != 2 else X[0])
    cnt = int((labels == lab).sum())
    assert G.delete_vector(lab) == P.delete(lab) == cnt
    assert G.index_size() == P.size() == n - cnt
    gl, gs = G.knn_batch(Q, 25)
    for i in range(nq):
        pl, ps, _ = P.topk(Q[i], 25)
        assert np.array_equal(gl[i], pl.astype(np.int64)) and np.array_equal(gs[i], ps)
    assert G.add_vector(X[7], lab) == 1
    assert dict(G.debug_info())["IS_MULTI_VALUE"] == 1
    G.close()
    P.close()


def test_multi_one_label_dominates(capi, port):
    """A label owning thousands of near vectors forces the selection to widen until k labels are found."""
    n, dim, k = 6000, 16, 20
    rng = np.random.default_rng(2)
    X = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    q = rng.uniform(-1, 1, dim).astype(np.float32)
    X[:4000] = q + 1e-3 * rng.standard_normal((4000, dim)).astype(np.float32)
    labels = np.concatenate([np.full(4000, 9), np.arange(100, 2100)]).astype(np.uint64)
    G, P = make_pair(capi, port, 0, dim, 0, X, labels)
    gl, gs = G.knn_query(q, k)
    pl, ps, _ = P.topk(q, k)
    assert gl[0][0] == 9
    assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps)
    G.close()
    P.close()


def test_delete_label_whose_id_list_holds_stale_entries(capi, port):
    """add X, A, Y, A; delete X; add A; delete Y — label A's id list is now [1, 0, 2]. Deleting A walks
    that list while rows move into the holes; the moved row's id must be patched from the BACK of its label's list
    (brute_force_multi.h:244-263), or a stale entry is rewritten, a row of A survives and the call reports 2 instead of 3."""
    dim = 8
    V = make_vectors(0, 5, dim, seed=2)
    G = capi.BFIndex(capi.BFParams(type=0, dim=dim, metric=0, multi=True, initialCapacity=0, blockSize=1024))
    P = port.PortIndex(0, dim, 0, multi=True)
    X, A, Y = 10, 20, 30
    for blob, lab in ((V[0], X), (V[1], A), (V[2], Y), (V[3], A)):
        G.add_vector(blob, lab)
        P.add(blob, lab)
    assert G.delete_vector(X) == P.delete(X) == 1
    G.add_vector(V[4], A)
    P.add(V[4], A)
    assert G.delete_vector(Y) == P.delete(Y) == 1
    assert G.index_size() == P.size() == 3
    assert G.delete_vector(A) == P.delete(A) == 3
    assert G.index_size() == P.size() == 0
    l, s = G.knn_query(V[1], 5)
    assert l.shape[1] == 0 or not (l == A).any()
    G.close()
    P.close()
