"""Parity tests proper: the CUDA path, called through the C-ABI (libvecsim_b200.so), against the oracle
port on the same seeded inputs and against the committed golden fixtures. Bit-exact everywhere:
labels, order and the fp scores themselves (the kernels reproduce the reference's AVX512 operation
order; tolerance stated by the north star would be 1e-5 rel, we assert equality)."""
import json
import os

import numpy as np
import pytest

from datagen import METRIC_NAMES, TYPE_NAMES, make_vectors, to_bf16

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
TYPES = {n: i for i, n in enumerate(TYPE_NAMES)}
METRICS = {n: i for i, n in enumerate(METRIC_NAMES)}
NPD = {"fp32": np.float32, "fp64": np.float64, "fp16": np.float16, "int8": np.int8, "uint8": np.uint8}


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    c.set_topk_mode(0)
    return c


def cast(tname, values):
    if tname == "bf16":
        return to_bf16(np.asarray(values, dtype=np.float32))
    return np.asarray(values, dtype=NPD[tname])


def make_pair(capi, port, vtype, dim, metric, X, labels=None, block_size=1024):
    G = capi.BFIndex(capi.BFParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=0,
                                   blockSize=block_size))
    P = port.PortIndex(vtype, dim, metric, block_size=block_size)
    if len(X):
        G.add_vectors(X, labels=labels)
        P.add_many(X, labels=labels)
    return G, P


DIMS = list(range(1, 70)) + [95, 96, 100, 127, 128, 129, 255, 256, 511, 512, 768, 1000, 1024, 1536]


@pytest.mark.parametrize("vtype", range(6), ids=TYPE_NAMES)
@pytest.mark.parametrize("metric", range(3), ids=METRIC_NAMES)
def test_scores_bit_exact_all_dims(capi, port, vtype, metric):
    """Every residual class of every dispatch tier: scan kernel (k = n) and gather kernel (ad-hoc)."""
    n, nq = 41, 3
    for dim in DIMS:
        X = make_vectors(vtype, n, dim, seed=dim * 31 + vtype)
        Q = make_vectors(vtype, nq, dim, seed=dim * 17 + metric + 5)
        if metric == 2:
            # a zero vector has no cosine (0/0 = NaN) and the reference's heap order is then unspecified
            X[(X == 0).all(1), 0] = 1
            Q[(Q == 0).all(1), 0] = 1
        G, P = make_pair(capi, port, vtype, dim, metric, X)
        gl, gs = G.knn_batch(Q, n)
        for i in range(nq):
            pl, ps, _ = P.topk(Q[i], n)
            assert np.array_equal(gl[i], pl.astype(np.int64)), (dim, i)
            assert np.array_equal(gs[i], ps), (dim, i, gs[i][:4], ps[:4])
            d = G.adhoc_distances(Q[i], np.arange(n))
            back = np.empty(n)
            back[pl.astype(np.int64)] = ps
            assert np.array_equal(d, back), (dim, i)
        G.close()
        P.close()


@pytest.mark.parametrize("kat", KAT["flat_topk"], ids=lambda k: k["src"].split(" ")[0].split("/")[-1])
def test_flat_topk_kat(capi, kat):
    """The reference's own unit-test expectations, through VecSimIndex_AddVector / _TopKQuery."""
    for tname in kat["types"]:
        for block_size in (1, 12, 1024):
            G = capi.BFIndex(capi.BFParams(type=TYPES[tname], dim=kat["dim"], metric=METRICS[kat["metric"]], multi=False,
                                           initialCapacity=0, blockSize=block_size))
            for i in range(kat["n"]):
                assert G.add_vector(cast(tname, [i] * kat["dim"]), i) == 1
            assert G.index_size() == kat["n"]
            q = cast(tname, [kat["query_value"]] * kat["dim"])
            order = capi.BY_ID if kat.get("order") == "BY_ID" else capi.BY_SCORE
            labels, scores = G.knn_query(q, kat["k"], order=order)
            assert G.last_code == 0
            if "expect_labels" in kat:
                assert labels[0].tolist() == kat["expect_labels"]
            if "expect_label_set" in kat:
                assert sorted(labels[0].tolist()) == kat["expect_label_set"]
            if "expect_scores" in kat:
                assert scores[0].tolist() == kat["expect_scores"]
            assert G.knn_query(q, 0)[0].size == 0
            G.close()


@pytest.mark.parametrize("vtype", range(6), ids=TYPE_NAMES)
@pytest.mark.parametrize("metric", range(3), ids=METRIC_NAMES)
def test_matches_captured_reference(capi, vtype, metric):
    """Outputs of the UNMODIFIED reference (tests/golden/make_golden.py), no oracle involved."""
    cases = np.load(os.path.join(HERE, "golden", "flat_cases.npz"))
    name = f"{TYPE_NAMES[vtype]}_{METRIC_NAMES[metric]}"
    X, Q = cases[name + "_X"], cases[name + "_Q"]
    G = capi.BFIndex(capi.BFParams(type=vtype, dim=X.shape[1], metric=metric, multi=False, initialCapacity=0, blockSize=64))
    G.add_vectors(X)
    k = cases[name + "_labels"].shape[1]
    for i in range(Q.shape[0]):
        labels, scores = G.knn_query(Q[i], k)
        assert np.array_equal(labels[0], cases[name + "_labels"][i])
        assert np.array_equal(scores[0], cases[name + "_scores"][i])
        rl, rs = G.range_query(Q[i], float(cases[name + "_radius"][i]))
        assert rl.shape[1] == cases[name + "_range_count"][i]
    gl, gs = G.knn_batch(Q, k)
    assert np.array_equal(gl, cases[name + "_labels"]) and np.array_equal(gs, cases[name + "_scores"])
    G.close()


def test_captured_ties(capi):
    cases = np.load(os.path.join(HERE, "golden", "flat_cases.npz"))
    G = capi.BFIndex(capi.BFParams(type=4, dim=3, metric=0, multi=False, initialCapacity=0, blockSize=16))
    G.add_vectors(cases["ties_X"], labels=cases["ties_lab"])
    labels, scores = G.knn_query(cases["ties_q"], len(cases["ties_labels"]))
    assert np.array_equal(labels[0].astype(np.uint64), cases["ties_labels"])
    assert np.array_equal(scores[0], cases["ties_scores"])
    G.close()


def test_ties_shuffled_labels_and_deletes(capi, port):
    """SURVEY App. A2 under the non-monotone regime (arbitrary labels, swap-deletes)."""
    rng = np.random.default_rng(7)
    for trial in range(60):
        n, dim, k = int(rng.integers(20, 400)), int(rng.integers(1, 4)), int(rng.integers(1, 40))
        X = rng.integers(-3, 4, (n, dim)).astype(np.int8)
        labels = rng.permutation(n * 3)[:n].astype(np.uint64)
        G, P = make_pair(capi, port, 4, dim, 0, X, labels=labels, block_size=7)
        for lab in labels[: n // 5]:
            assert G.delete_vector(int(lab)) == P.delete(int(lab)) == 1
        assert G.delete_vector(10 ** 9) == 0
        assert G.index_size() == P.size()
        Q = rng.integers(-3, 4, (3, dim)).astype(np.int8)
        gl, gs = G.knn_batch(Q, k)
        for i in range(3):
            pl, ps, _ = P.topk(Q[i], k)
            m = len(pl)
            assert np.array_equal(gl[i][:m], pl.astype(np.int64)), trial
            assert np.array_equal(gs[i][:m], ps), trial
            assert (gl[i][m:] == -1).all()
        G.close()
        P.close()


@pytest.mark.parametrize("vtype,metric,dim,n,k,nq", [
    (0, 0, 128, 20000, 10, 1),      # BASELINE config 1 shape (reduced n)
    (0, 1, 768, 6000, 100, 33),     # config 2 shape: fp32 IP d=768 K=100, several query chunks
    (4, 2, 512, 8000, 10, 40),      # config 3 shape: int8 cosine d=512
    (2, 1, 1024, 4000, 100, 20),    # config 4 shape: bf16 IP d=1024
    (1, 0, 96, 5000, 50, 9),        # fp64
    (3, 2, 200, 5000, 20, 5),       # fp16 cosine with a residual
    (5, 0, 300, 5000, 20, 17),      # uint8 L2, ragged dim
])
def test_batch_topk_vs_oracle(capi, port, vtype, metric, dim, n, k, nq):
    dist = "normal" if metric == 1 and vtype < 4 else "uniform"
    X = make_vectors(vtype, n, dim, seed=n + dim, dist=dist)
    Q = make_vectors(vtype, nq, dim, seed=n + dim + 1, dist=dist)
    G, P = make_pair(capi, port, vtype, dim, metric, X)
    gl, gs = G.knn_batch(Q, k)
    stats = G.last_query_stats()
    assert stats["kernel_launches"] > 0
    for i in range(nq):
        pl, ps, _ = P.topk(Q[i], k)
        assert np.array_equal(gl[i], pl.astype(np.int64)), i
        assert np.array_equal(gs[i], ps), i
    # the single-query API and the reply-object batch API give the same answers
    l1, s1 = G.knn_query(Q[0], k)
    assert np.array_equal(l1[0], gl[0]) and np.array_equal(s1[0], gs[0])
    reps = G.knn_batch_replies(Q[:3], k, order=capi.BY_ID)
    for i, (rl, rs, code) in enumerate(reps):
        assert code == 0 and np.array_equal(np.sort(gl[i]), rl)
    G.close()
    P.close()


def test_large_k_host_sort_and_k_over_n(capi, port):
    n, dim = 6000, 16
    X = make_vectors(0, n, dim, seed=1, dist="grid")
    G, P = make_pair(capi, port, 0, dim, 0, X)
    q = make_vectors(0, 1, dim, seed=2, dist="grid")[0]
    for k in (5000, n, n + 10):
        gl, gs = G.knn_query(q, k)
        pl, ps, _ = P.topk(q, k)
        assert np.array_equal(gl[0], pl.astype(np.int64))
        assert np.array_equal(gs[0], ps)
    G.close()
    P.close()


def test_range_query_vs_oracle(capi, port):
    for vtype, metric, dim in [(0, 0, 32), (0, 2, 50), (4, 0, 64), (1, 1, 24), (2, 0, 40)]:
        n = 3000
        X = make_vectors(vtype, n, dim, seed=dim)
        G, P = make_pair(capi, port, vtype, dim, metric, X)
        q = make_vectors(vtype, 1, dim, seed=dim + 9)[0]
        _, allscores, _ = P.topk(q, n)
        for frac in (0.0, 0.01, 0.3, 3.0):
            radius = float(allscores[min(int(n * frac), n - 1)]) if frac <= 1 else float(allscores[-1]) + 1
            if radius < 0:
                continue
            for order in (capi.BY_SCORE, capi.BY_ID):
                gl, gs = G.range_query(q, radius, order=order)
                pl, ps, _ = P.range(q, radius, order)
                assert np.array_equal(gs[0], ps)
                if order == capi.BY_ID:
                    assert np.array_equal(gl[0], pl.astype(np.int64))
                else:
                    assert sorted(gl[0].tolist()) == sorted(pl.tolist())
        G.close()
        P.close()


def test_batch_iterator_vs_oracle(capi, port):
    """tests/unit/test_bruteforce.cpp:959-1055: batches come back in score order, no repeats,
    non-unique scores allowed to permute inside a tie group."""
    n, dim = 1000, 8
    X = make_vectors(0, n, dim, seed=3, dist="grid")
    G, P = make_pair(capi, port, 0, dim, 0, X)
    q = make_vectors(0, 1, dim, seed=4, dist="grid")[0]
    gi, pi = G.create_batch_iterator(q), P.batch_iterator(q)
    seen = []
    while gi.has_next():
        assert pi.has_next()
        gl, gs = gi.get_next_results(64)
        pl, ps, _ = pi.next(64)
        assert np.array_equal(gs[0], ps)
        seen += gl[0].tolist()
    assert not pi.has_next()
    assert sorted(seen) == list(range(n))
    gi.reset()
    gl, gs = gi.get_next_results(10, order=capi.BY_ID)
    assert np.array_equal(gl[0], np.sort(gl[0]))
    gi.close()
    pi.close()
    G.close()
    P.close()


def test_update_delete_distance_from(capi, port):
    dim = 20
    X = make_vectors(0, 200, dim, seed=8)
    G, P = make_pair(capi, port, 0, dim, 0, X)
    newv = make_vectors(0, 1, dim, seed=9)[0]
    assert G.add_vector(newv, 17) == 0 and P.add(newv, 17) == 0       # overwrite (L2: no preprocessing involved)
    assert G.index_size() == 200
    q = make_vectors(0, 1, dim, seed=10)[0]
    assert G.get_distance_from(17, q) == P.distance_from(17, q)
    assert np.isnan(G.get_distance_from(9999, q))
    gl, gs = G.knn_query(q, 200)
    pl, ps, _ = P.topk(q, 200)
    assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps)
    d = G.adhoc_distances(q, [3, 9999, 17])
    assert d[0] == P.distance_from(3, q) and np.isnan(d[1]) and d[2] == P.distance_from(17, q)
    G.close()
    P.close()


def test_empty_index_inf_and_timeout(capi, port):
    G = capi.BFIndex(capi.BFParams(type=0, dim=4, metric=0, multi=False, initialCapacity=0, blockSize=5))
    q = np.ones(4, dtype=np.float32)
    l, s = G.knn_query(q, 5)                      # test_bruteforce.cpp:814-862
    assert l.size == 0 and G.last_code == 0
    l, s = G.range_query(q, 1.0)
    assert l.size == 0
    it = G.create_batch_iterator(q)
    assert not it.has_next()
    it.close()
    # inf scores (test_bruteforce.cpp:864-900)
    big = np.finfo(np.float32).max
    for i in range(5):
        G.add_vector(np.full(4, float(i), dtype=np.float32), i)
    G.add_vector(np.full(4, big, dtype=np.float32), 100)
    G.add_vector(np.full(4, -big, dtype=np.float32), 101)
    l, s = G.knn_query(q, 7)
    assert l[0][:5].tolist() == [1, 0, 2, 3, 4] and set(l[0][5:].tolist()) == {100, 101}
    assert np.isinf(s[0][5:]).all()
    # timeouts (test_bruteforce.cpp:1489-1565): top-k empty + TimedOut, range partial + TimedOut
    capi.set_timeout_callback(lambda ctx: 1)
    try:
        l, s = G.knn_query(q, 3)
        assert l.size == 0 and G.last_code == capi.VecSim_QueryReply_TimedOut
        l, s = G.range_query(q, 10.0)
        assert G.last_code == capi.VecSim_QueryReply_TimedOut
        it = G.create_batch_iterator(q)
        l, s = it.get_next_results(2)
        assert it.last_code == capi.VecSim_QueryReply_TimedOut and l.size == 0
        it.close()
    finally:
        capi.set_timeout_callback(None)
    l, s = G.knn_query(q, 3)
    assert l.shape[1] == 3 and G.last_code == 0
    G.close()


def test_against_reference_if_present(capi, ref):
    """When oracle/_ref travelled to this box and the host has the modelled AVX512 tier: the product
    against the unmodified reference directly."""
    feats = set(ref.host_features())
    if not {"avx512f", "avx512bw", "avx512vl", "avx512vnni", "avx512vbmi2", "avx512_bf16"} <= feats:
        pytest.skip("host CPU dispatches a different tier than the one the kernels reproduce")
    ref.set_disabled_features("avx512_fp16")
    try:
        for vtype, metric, dim in [(0, 1, 768), (0, 0, 100), (2, 1, 1024), (3, 0, 72), (4, 2, 512), (1, 2, 33)]:
            X = make_vectors(vtype, 3000, dim, seed=dim)
            Q = make_vectors(vtype, 4, dim, seed=dim + 1)
            G = capi.BFIndex(capi.BFParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=0, blockSize=0))
            G.add_vectors(X)
            R = ref.RefIndex(vtype, dim, metric)
            R.add_many(X)
            gl, gs = G.knn_batch(Q, 50)
            for i in range(4):
                rl, rs, _ = R.topk(Q[i], 50)
                assert np.array_equal(gl[i], rl.astype(np.int64))
                assert np.array_equal(gs[i], rs)
            G.close()
            R.close()
    finally:
        ref.set_disabled_features()


def test_debug_info_iterator_flat(capi):
    """VecSimIndex_DebugInfoIterator: names and order of brute_force.h:348-365 + vec_sim_index.h:268-310."""
    G = capi.BFIndex(capi.BFParams(type=2, dim=40, metric=2, multi=False, initialCapacity=0, blockSize=77))
    G.add_vectors(make_vectors(2, 10, 40, seed=1))
    info = G.debug_info()
    assert [k for k, _ in info] == ["ALGORITHM", "TYPE", "DIMENSION", "METRIC", "IS_MULTI_VALUE", "IS_DISK", "INDEX_SIZE",
                                    "INDEX_LABEL_COUNT", "MEMORY", "LAST_SEARCH_MODE", "BLOCK_SIZE"]
    d = dict(info)
    assert (d["ALGORITHM"], d["TYPE"], d["METRIC"], d["DIMENSION"], d["INDEX_SIZE"], d["BLOCK_SIZE"]) == \
        ("FLAT", "BFLOAT16", "COSINE", 40, 10, 77)
    rc = capi.lib().VecSimDebug_GetElementNeighborsInHNSWGraph  # flat index -> BadIndex
    import ctypes as C
    out = C.POINTER(C.POINTER(C.c_int))()
    assert rc(G._h, 0, C.byref(out)) == 1
    G.close()


def test_concurrent_single_query_callers_are_combined(capi, port):
    """SURVEY §8b "Threading": RediSearch calls VecSimIndex_TopKQuery from its worker threads concurrently. Callers that
    arrive while a device call is running are served together by the next one (flat combining in FlatIndex::topKQuery);
    every caller gets exactly the reply it would get alone — also with different k per caller."""
    import threading
    n, dim = 30000, 64
    X = make_vectors(0, n, dim, seed=71)
    Q = make_vectors(0, 64, dim, seed=72)
    G = capi.BFIndex(capi.BFParams(type=0, dim=dim, metric=0, multi=False, initialCapacity=n, blockSize=1024))
    G.add_vectors(X)
    P = port.PortIndex(0, dim, 0)
    P.add_many(X)
    want = {}
    for i in range(len(Q)):
        k = 5 + (i % 4) * 7
        l, s, _ = P.topk(Q[i], k)
        want[i] = (l.astype(np.int64), s)
    errors = []

    def worker(t):
        try:
            for rep in range(6):
                for i in range(t, len(Q), 8):
                    k = 5 + (i % 4) * 7
                    l, s = G.knn_query(Q[i], k)
                    if not (np.array_equal(l[0], want[i][0]) and np.array_equal(s[0], want[i][1])):
                        errors.append((t, i))
        except Exception as e:  # pragma: no cover
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:5]
    G.close()
    P.close()
