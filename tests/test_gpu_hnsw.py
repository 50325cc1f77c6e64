"""HNSW parity on the device, through the C-ABI (libvecsim_b200.so): traversal results on the
reference's own graph, and the device builder against the graph the reference builds
(SURVEY §8 rows a13-a15). Everything is compared bit-exact: labels, order, fp scores, link lists."""
import os

import numpy as np
import pytest

from datagen import METRIC_NAMES, TYPE_NAMES, make_vectors

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "hnsw_case.npz"))
N, DIM, M, EFC, NQ, K = 1500, 24, 8, 48, 16, 10  # tests/golden/make_hnsw_golden.py


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    return c


def new_index(capi, vtype, dim, metric, M=16, efc=200, ef=10):
    return capi.HNSWIndex(capi.HNSWParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=0,
                                          blockSize=1024, M=M, efConstruction=efc, efRuntime=ef, epsilon=0.01))


def gold_graph(name):
    levels = GOLD[name + "_levels"]
    nl = int(GOLD[name + "_entry"][1]) + 1
    links = [GOLD[f"{name}_links{l}"] for l in range(nl)]
    counts = [GOLD[f"{name}_counts{l}"] for l in range(nl)]
    return levels, links, counts, int(GOLD[name + "_entry"][0]), int(GOLD[name + "_entry"][1])


def graph_records(levels, links, counts, M):
    """(l0 [n, 2M+1], upper [records, M+1]) in the export layout of include/vsgpu.h."""
    n = len(levels)
    l0 = np.zeros((n, 2 * M + 1), dtype=np.uint32)
    l0[:, 0] = counts[0]
    l0[:, 1:] = np.where(np.arange(2 * M)[None, :] < counts[0][:, None], links[0], 0)
    recs = []
    for i in np.nonzero(levels)[0]:
        for lvl in range(1, int(levels[i]) + 1):
            r = np.zeros(M + 1, dtype=np.uint32)
            c = int(counts[lvl][i])
            r[0] = c
            r[1:1 + c] = links[lvl][i][:c]
            recs.append(r)
    return l0, (np.stack(recs) if recs else np.zeros((0, M + 1), dtype=np.uint32))


def masked(rec):
    """zero the unused tail of each link record so records compare as arrays"""
    rec = rec.copy()
    w = rec.shape[1] - 1
    rec[:, 1:] = np.where(np.arange(w)[None, :] < rec[:, :1], rec[:, 1:], 0)
    return rec


@pytest.mark.parametrize("name,metric", [("l2", 0), ("cos", 2)])
def test_traversal_on_reference_graph_matches_golden(capi, name, metric):
    """Import the graph the reference built, search it on the device: top-k and range results are the
    reference's, bit for bit."""
    levels, links, counts, entry, maxl = gold_graph(name)
    G = new_index(capi, 0, DIM, metric, M=M, efc=EFC)
    G.import_graph(GOLD[name + "_stored"], levels, links, counts, entry, maxl, processed=True)
    assert G.index_size() == N
    Q = GOLD[name + "_Q"]
    for ef in (10, 40):
        G.set_ef(ef)
        labels, scores = G.knn_batch(Q, K)
        assert np.array_equal(labels, GOLD[f"{name}_labels_ef{ef}"]), ef
        assert np.array_equal(scores, GOLD[f"{name}_scores_ef{ef}"]), ef
        # single-query API gives the same as the batch
        l1, s1 = G.knn_query(Q[3], K)
        assert np.array_equal(l1[0], labels[3]) and np.array_equal(s1[0], scores[3])
    off = 0
    for i in range(NQ):
        n = int(GOLD[name + "_range_n"][i])
        l, s = G.range_query(Q[i], float(GOLD[name + "_radius"][i]))
        assert l.shape[1] == n
        assert np.array_equal(l[0], GOLD[name + "_range_labels"][off:off + n])
        assert np.array_equal(s[0], GOLD[name + "_range_scores"][off:off + n])
        off += n
    st = G.hnsw_stats()
    assert st["dist_evals"] > 0 and st["hops"] > 0
    G.close()


@pytest.mark.parametrize("name,metric", [("l2", 0), ("cos", 2)])
def test_builder_reproduces_reference_graph(capi, name, metric):
    """VecSimIndex_AddVector on the device: same levels, entry point and link lists (in order) as the
    reference's single-threaded build over the same vectors."""
    levels, links, counts, entry, maxl = gold_graph(name)
    G = new_index(capi, 0, DIM, metric, M=M, efc=EFC)
    X = GOLD[name + "_X"]
    assert G.add_vectors(X[:700]) == 700
    for i in range(700, 720):  # one at a time through VecSimIndex_AddVector
        assert G.add_vector(X[i], i) == 1
    assert G.add_vectors(X[720:], first_label=720) == N - 720
    g = G.export_graph(N)
    assert np.array_equal(g["levels"], levels)
    assert (g["entry"], g["max_level"]) == (entry, maxl)
    l0, upper = graph_records(levels, links, counts, M)
    assert np.array_equal(masked(g["l0"])[:, 0], l0[:, 0]), "level-0 degrees differ"
    assert np.array_equal(masked(g["l0"]), l0), "level-0 links differ"
    assert np.array_equal(masked(g["upper"]), upper), "upper-level links differ"
    Q = GOLD[name + "_Q"]
    G.set_ef(40)
    labels, scores = G.knn_batch(Q, K)
    assert np.array_equal(labels, GOLD[f"{name}_labels_ef40"])
    assert np.array_equal(scores, GOLD[f"{name}_scores_ef40"])
    G.close()


CASES = [(0, 1, 128), (1, 0, 20), (2, 1, 64), (2, 0, 40), (3, 2, 48), (4, 2, 64), (4, 0, 33), (5, 1, 100), (0, 0, 5),
         (3, 0, 12), (0, 2, 768)]


@pytest.mark.parametrize("vtype,metric,dim", CASES, ids=lambda v: str(v))
def test_build_and_search_match_oracle(capi, port, vtype, metric, dim):
    """Every distance policy (chain fp32/fp64/bf16/fp16, integer, scalar tiers): build on the device and in
    the oracle (pinned to the reference by tests/test_oracle_vs_reference.py) from the same vectors; the
    graphs and the query results must be identical."""
    port.set_tier(port.TIER_AVX512)
    n, nq, k, M = 1200, 12, 8, 6
    X = make_vectors(vtype, n, dim, seed=900 + vtype * 7 + metric)
    Q = make_vectors(vtype, nq, dim, seed=901 + vtype * 7 + metric)
    if metric == 2:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    P = port.PortHnsw(vtype, dim, metric, M=M, ef_construction=40, ef_runtime=10)
    P.add_many(X)
    G = new_index(capi, vtype, dim, metric, M=M, efc=40)
    assert G.add_vectors(X) == n
    gp, gg = P.export(), G.export_graph(n)
    assert np.array_equal(gg["levels"], gp["levels"])
    assert (gg["entry"], gg["max_level"]) == (gp["entry"], gp["max_level"])
    l0, upper = graph_records(gp["levels"], gp["links"], gp["counts"], M)
    assert np.array_equal(masked(gg["l0"]), l0), "level-0 links differ"
    assert np.array_equal(masked(gg["upper"]), upper), "upper-level links differ"
    for ef in (10, 64):
        G.set_ef(ef)
        labels, scores = G.knn_batch(Q, k)
        for i in range(nq):
            pl, ps, _ = P.topk(Q[i], k, ef_runtime=ef)
            assert np.array_equal(labels[i], pl.astype(np.int64)), (TYPE_NAMES[vtype], METRIC_NAMES[metric], ef, i)
            assert np.array_equal(scores[i], ps), (TYPE_NAMES[vtype], METRIC_NAMES[metric], ef, i)
    for i in range(nq):
        radius = max(float(P.topk(Q[i], k, ef_runtime=64)[1][4]), 0.0)
        pl, ps, _ = P.range(Q[i], radius)
        gl, gs = G.range_query(Q[i], radius)
        assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps), (vtype, metric, i)
    G.close()
    P.close()


@pytest.mark.parametrize("vtype,metric,dim", CASES[:8], ids=lambda v: str(v))
def test_build_and_search_match_live_reference(capi, ref, vtype, metric, dim):
    """Same, against the unmodified reference itself (oracle/_ref) where the host CPU dispatches the tier the
    device kernels reproduce."""
    feats = ref.host_features()
    if vtype == 2 and metric != 0 and "avx512_bf16" not in feats:
        pytest.skip("host lacks avx512_bf16: the reference dispatches a different bf16 IP tier here")
    if not all(f in feats for f in ("avx512f", "avx512bw", "avx512vl", "avx512vnni", "avx512vbmi2")):
        pytest.skip("host lacks the AVX512 tier")
    ref.set_disabled_features("avx512_fp16")
    n, nq, k = 1200, 12, 8
    X = make_vectors(vtype, n, dim, seed=900 + vtype * 7 + metric)
    Q = make_vectors(vtype, nq, dim, seed=901 + vtype * 7 + metric)
    if metric == 2:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    R = ref.RefIndex(vtype, dim, metric, algo="hnsw", M=6, ef_construction=40, ef_runtime=10)
    R.add_many(X)
    G = new_index(capi, vtype, dim, metric, M=6, efc=40)
    assert G.add_vectors(X) == n
    for ef in (10, 64):
        G.set_ef(ef)
        labels, scores = G.knn_batch(Q, k)
        for i in range(nq):
            rl, rs, code = R.topk(Q[i], k, ef_runtime=ef)
            assert code == 0
            assert np.array_equal(labels[i], rl.astype(np.int64)), (TYPE_NAMES[vtype], METRIC_NAMES[metric], ef, i)
            assert np.array_equal(scores[i], rs), (TYPE_NAMES[vtype], METRIC_NAMES[metric], ef, i)
    G.close()
    R.close()
    ref.set_disabled_features()


def test_deleted_nodes_are_traversed_not_returned(capi):
    n, dim = 600, 16
    X = make_vectors(0, n, dim, seed=5)
    G = new_index(capi, 0, dim, 0, M=8, efc=40)
    G.add_vectors(X)
    G.set_ef(50)
    l0, s0 = G.knn_query(X[17], 5)
    assert l0[0][0] == 17 and s0[0][0] == 0.0
    assert G.delete_vector(17) == 1 and G.delete_vector(17) == 0
    assert G.index_size() == n - 1
    l1, _ = G.knn_query(X[17], 5)
    assert 17 not in l1[0]
    assert list(l1[0][:4]) == list(l0[0][1:5])
    # overwrite = tombstone + append
    assert G.add_vector(X[18] * 0.5, 18) == 0
    l2, _ = G.knn_query(X[18] * 0.5, 1)
    assert l2[0][0] == 18
    G.close()


def test_empty_and_small(capi):
    G = new_index(capi, 0, 8, 0, M=4, efc=10)
    q = np.ones(8, dtype=np.float32)
    l, s = G.knn_query(q, 5)
    assert l.shape == (1, 0)
    l, s = G.range_query(q, 10.0)
    assert l.shape == (1, 0)
    G.add_vector(q * 2, 42)
    l, s = G.knn_query(q, 5)
    assert list(l[0]) == [42] and s[0][0] == 8.0
    G.add_vector(q * 3, 7)
    l, s = G.knn_query(q, 5)
    assert list(l[0]) == [42, 7]
    l, s = G.range_query(q, 8.0)
    assert list(l[0]) == [42]
    it = G.create_batch_iterator(q)
    assert it.has_next()
    l, _ = it.get_next_results(1)
    assert list(l[0]) == [42]
    l, _ = it.get_next_results(5)
    assert list(l[0]) == [7]
    assert not it.has_next()
    it.close()
    G.close()


@pytest.mark.parametrize("vtype,metric,dim", [(0, 0, 16), (0, 2, 32), (4, 2, 64), (1, 1, 12)], ids=lambda v: str(v))
def test_batch_iterator_matches_oracle(capi, port, vtype, metric, dim):
    """VecSimBatchIterator on an HNSW index: the resumable device traversal returns the same batches as the
    reference's HNSW_BatchIterator (oracle pinned by test_hnsw_port_batch_iterator_matches_reference)."""
    port.set_tier(port.TIER_AVX512)
    n = 800
    X = make_vectors(vtype, n, dim, seed=71 + metric)
    Q = make_vectors(vtype, 3, dim, seed=72 + metric)
    if metric == 2:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    P = port.PortHnsw(vtype, dim, metric, M=6, ef_construction=40, ef_runtime=10)
    P.add_many(X)
    G = new_index(capi, vtype, dim, metric, M=6, efc=40, ef=10)
    G.add_vectors(X)
    for q in Q:
        for sched in ([5, 5, 5, 20, 1, 100], [1, 2, 3], [50, 50], [1000]):
            gi, pi = G.create_batch_iterator(q), P.batch_iterator(q)
            for rounds in range(2):
                for nres in sched:
                    assert bool(gi.has_next()) == pi.has_next()
                    gl, gs = gi.get_next_results(nres)
                    pl, ps, _ = pi.next(nres)
                    assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps), (sched, nres)
                assert bool(gi.has_next()) == pi.has_next()
                gi.reset()
                pi.reset()
            gi.close()
            pi.close()
    # BY_ID order + efRuntime from the query params
    qp = capi.VecSimQueryParams()
    qp.hnswRuntimeParams.efRuntime = 30
    gi, pi = G.create_batch_iterator(Q[0], qp), P.batch_iterator(Q[0], ef_runtime=30)
    gl, gs = gi.get_next_results(12, capi.BY_ID)
    pl, ps, _ = pi.next(12, port.BY_ID)
    assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps)
    gi.close()
    pi.close()
    G.close()
    P.close()


def test_builder_cta_wide_revisit_path(capi, monkeypatch):
    """The builder's fallback when shared memory has no room for per-warp revisit scratch (large M): same graph."""
    levels, links, counts, entry, maxl = gold_graph("l2")
    monkeypatch.setenv("VSGPU_HNSW_RV_WARPS", "0")
    G = new_index(capi, 0, DIM, 0, M=M, efc=EFC)
    G.add_vectors(GOLD["l2_X"])
    g = G.export_graph(N)
    l0, upper = graph_records(levels, links, counts, M)
    assert np.array_equal(g["levels"], levels) and (g["entry"], g["max_level"]) == (entry, maxl)
    assert np.array_equal(masked(g["l0"]), l0) and np.array_equal(masked(g["upper"]), upper)
    G.close()


def test_debug_info_iterator_and_element_neighbors(capi, port):
    """VecSimIndex_DebugInfoIterator field names/order (hnsw.h:2217-2272) and
    VecSimDebug_GetElementNeighborsInHNSWGraph (hnsw.h:2414-2441) against the oracle's graph."""
    n, dim, Mv = 500, 16, 5
    X = make_vectors(0, n, dim, seed=9)
    labels = np.arange(n, dtype=np.uint64) * 3 + 7          # labels != internal ids
    G = new_index(capi, 0, dim, 0, M=Mv, efc=30, ef=12)
    G.add_vectors(X, labels=labels)
    P = port.PortHnsw(0, dim, 0, M=Mv, ef_construction=30, ef_runtime=12)
    P.add_many(X, labels=labels)
    G.knn_query(X[0], 3)
    info = G.debug_info()
    assert [k for k, _ in info] == ["ALGORITHM", "TYPE", "DIMENSION", "METRIC", "IS_MULTI_VALUE", "IS_DISK", "INDEX_SIZE",
                                    "INDEX_LABEL_COUNT", "MEMORY", "LAST_SEARCH_MODE", "BLOCK_SIZE", "M", "EF_CONSTRUCTION",
                                    "EF_RUNTIME", "MAX_LEVEL", "ENTRYPOINT", "EPSILON", "NUMBER_OF_MARKED_DELETED"]
    d = dict(info)
    gp = P.export()
    assert d["ALGORITHM"] == "HNSW" and d["TYPE"] == "FLOAT32" and d["METRIC"] == "L2" and d["DIMENSION"] == dim
    assert d["INDEX_SIZE"] == n and d["M"] == Mv and d["EF_CONSTRUCTION"] == 30 and d["EF_RUNTIME"] == 12
    assert d["MAX_LEVEL"] == gp["max_level"] and d["ENTRYPOINT"] == int(labels[gp["entry"]])
    assert d["LAST_SEARCH_MODE"] == "STANDARD_KNN" and d["EPSILON"] == 0.01
    for node in (0, 17, gp["entry"], n - 1):
        rc, levels = G.element_neighbors(int(labels[node]))
        assert rc == 0 and len(levels) == int(gp["levels"][node]) + 1
        for lvl, got in enumerate(levels):
            c = int(gp["counts"][lvl][node])
            assert got == [int(labels[i]) for i in gp["links"][lvl][node][:c]]
    assert G.element_neighbors(1)[0] == 2                    # VecSimDebugCommandCode_LabelNotExists
    G.close()
    P.close()


def test_batched_builder_equals_sequential_builder(capi, monkeypatch):
    """The builder searches for up to one element per SM concurrently and commits them in order while their read sets
    are untouched (csrc/vsgpu_hnsw.cu, mode 1 / mode 2). The graph must be the sequential one, link for link."""
    n, dim, Mv = 9000, 24, 8
    X = make_vectors(0, n, dim, seed=123)
    monkeypatch.setenv("VSGPU_HNSW_SEQ_BUILD", "1")
    A = new_index(capi, 0, dim, 0, M=Mv, efc=60)
    A.add_vectors(X)
    ga = A.export_graph(n)
    monkeypatch.delenv("VSGPU_HNSW_SEQ_BUILD")
    monkeypatch.setenv("VSGPU_HNSW_BATCH_BUILD", "1")
    B = new_index(capi, 0, dim, 0, M=Mv, efc=60)
    B.add_vectors(X[:5000])
    B.add_vectors(X[5000:], first_label=5000)
    gb = B.export_graph(n)
    assert np.array_equal(ga["levels"], gb["levels"]) and (ga["entry"], ga["max_level"]) == (gb["entry"], gb["max_level"])
    assert np.array_equal(masked(ga["l0"]), masked(gb["l0"]))
    assert np.array_equal(masked(ga["upper"]), masked(gb["upper"]))
    A.close()
    B.close()
