"""Pins the C restatement (oracle/vs_oracle.c) against the UNMODIFIED reference (oracle/_ref),
bit for bit. Mirrors the reference's own strategy of walking the dispatcher down tier by tier
(tests/unit/test_spaces.cpp:703-709) by masking CPU feature bits."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))

from datagen import (BFLOAT16, COSINE, FLOAT16, FLOAT32, FLOAT64, INT8, IP, L2, METRIC_NAMES,
                     TYPE_NAMES, UINT8, make_vectors)

ALL_SIMD = ["sse", "sse3", "sse4_1", "avx", "avx2", "fma3", "f16c", "avx512f", "avx512bw",
            "avx512vl", "avx512vnni", "avx512vbmi2", "avx512_bf16", "avx512_fp16"]
NEEDED = {"avx512f", "avx512bw", "avx512vl", "avx512vnni", "avx512vbmi2", "fma3", "f16c"}


def _processed(port, vtype, metric, dim, raw):
    """raw caller blob -> processed blob as the index would store/query it."""
    if metric == COSINE:
        if vtype in (INT8, UINT8):
            b = np.zeros(dim + 4, dtype=np.uint8)
            b[:dim] = raw.view(np.uint8)
            port.normalize(vtype, dim, b)
            return b
        b = raw.copy()
        port.normalize(vtype, dim, b)
        return b
    return raw


DIMS = list(range(1, 100)) + [127, 128, 129, 255, 256, 512, 768, 777, 1000, 1024]


@pytest.mark.parametrize("vtype", range(6), ids=TYPE_NAMES)
@pytest.mark.parametrize("metric", range(3), ids=METRIC_NAMES)
@pytest.mark.parametrize("tier", ["avx512", "avx512_nobf16", "naive"])
def test_distance_bit_exact(ref, port, vtype, metric, tier):
    feats = set(ref.host_features())
    if tier != "naive" and not NEEDED <= feats:
        pytest.skip("host lacks the AVX512 feature set the restated tier models")
    if tier == "avx512":
        if vtype == BFLOAT16 and "avx512_bf16" not in feats:
            pytest.skip("no avx512_bf16 on this host")
        ref.set_disabled_features("avx512_fp16")
        port.set_tier(port.TIER_AVX512)
    elif tier == "avx512_nobf16":
        ref.set_disabled_features("avx512_fp16", "avx512_bf16")
        port.set_tier(port.TIER_AVX512_NOBF16)
    else:
        ref.set_disabled_features(*ALL_SIMD)
        port.set_tier(port.TIER_NAIVE)
    try:
        bad = []
        for dim in DIMS:
            A = make_vectors(vtype, 4, dim, seed=dim * 7 + vtype)
            B = make_vectors(vtype, 4, dim, seed=dim * 13 + metric + 1000)
            for i in range(4):
                a = _processed(port, vtype, metric, dim, A[i])
                b = _processed(port, vtype, metric, dim, B[i])
                r = ref.distance(vtype, metric, a, b, dim)
                p = port.distance(vtype, metric, a, b, dim)
                if not (r == p or (np.isnan(r) and np.isnan(p))):
                    bad.append((dim, r, p))
        assert not bad, bad[:5]
    finally:
        ref.set_disabled_features()
        port.set_tier(port.TIER_AVX512)


@pytest.mark.parametrize("vtype", range(6), ids=TYPE_NAMES)
def test_normalize_bit_exact(ref, port, vtype):
    for dim in [1, 3, 4, 17, 128, 513]:
        raw = make_vectors(vtype, 3, dim, seed=dim + 5)
        for i in range(3):
            if vtype in (INT8, UINT8):
                a = np.zeros(dim + 4, dtype=np.uint8)
                a[:dim] = raw[i].view(np.uint8)
            else:
                a = raw[i].copy()
            b = a.copy()
            ref.normalize(vtype, dim, a)
            port.normalize(vtype, dim, b)
            assert a.tobytes() == b.tobytes()


def _flat_pair(ref, port, vtype, dim, metric, n, seed, block_size=1024, labels=None, dist="uniform"):
    X = make_vectors(vtype, n, dim, seed, dist)
    R = ref.RefIndex(vtype, dim, metric, block_size=block_size)
    P = port.PortIndex(vtype, dim, metric, block_size=block_size)
    R.add_many(X, labels=labels)
    P.add_many(X, labels=labels)
    return X, R, P


@pytest.mark.parametrize("vtype", range(6), ids=TYPE_NAMES)
@pytest.mark.parametrize("metric", range(3), ids=METRIC_NAMES)
def test_flat_topk_range_identical(ref, port, vtype, metric):
    feats = set(ref.host_features())
    if not NEEDED <= feats or (vtype == BFLOAT16 and "avx512_bf16" not in feats):
        pytest.skip("host lacks the modelled AVX512 tier")
    ref.set_disabled_features("avx512_fp16")
    try:
        for dim, n, k in [(4, 300, 11), (33, 500, 10), (128, 1000, 100)]:
            X, R, P = _flat_pair(ref, port, vtype, dim, metric, n, seed=n + dim)
            Q = make_vectors(vtype, 5, dim, seed=99 + dim)
            for q in Q:
                for order in (0, 1):
                    rl, rs, rc = R.topk(q, k, order)
                    pl, ps, pc = P.topk(q, k, order)
                    assert rc == pc == 0
                    assert np.array_equal(rl, pl)
                    assert np.array_equal(rs, ps)
                # a radius that catches ~5% of the rows
                _, allscores, _ = P.topk(q, n)
                radius = float(allscores[n // 20])
                if radius >= 0:
                    rl, rs, _ = R.range(q, radius, 1)
                    pl, ps, _ = P.range(q, radius, 1)
                    assert np.array_equal(rl, pl) and np.array_equal(rs, ps)
                    rl, rs, _ = R.range(q, radius, 0)
                    pl, ps, _ = P.range(q, radius, 0)
                    assert np.array_equal(rs, ps)
                    assert sorted(rl.tolist()) == sorted(pl.tolist())
            R.close()
            P.close()
    finally:
        ref.set_disabled_features()


def test_flat_ties_shuffled_labels(ref, port):
    """int8 L2 on tiny dims gives surplus ties at the k-th score; with shuffled
    labels the result depends on scan order AND labels. 200 random trials."""
    rng = np.random.default_rng(7)
    for trial in range(200):
        n, dim, k = int(rng.integers(20, 120)), int(rng.integers(1, 4)), int(rng.integers(1, 30))
        X = rng.integers(-3, 4, (n, dim)).astype(np.int8)
        labels = rng.permutation(n * 3)[:n].astype(np.uint64)
        R = ref.RefIndex(INT8, dim, L2, block_size=7)
        P = port.PortIndex(INT8, dim, L2, block_size=7)
        R.add_many(X, labels=labels)
        P.add_many(X, labels=labels)
        # a few deletes: swap-with-last changes the scan order
        for lab in labels[: n // 5]:
            assert R.delete(int(lab)) == P.delete(int(lab)) == 1
        q = rng.integers(-3, 4, dim).astype(np.int8)
        rl, rs, _ = R.topk(q, k)
        pl, ps, _ = P.topk(q, k)
        assert np.array_equal(rl, pl), trial
        assert np.array_equal(rs, ps), trial
        R.close()
        P.close()


def test_batch_iterator_matches_up_to_ties(ref, port):
    n, dim = 500, 8
    X, R, P = _flat_pair(ref, port, FLOAT32, dim, L2, n, seed=3, dist="grid")
    q = make_vectors(FLOAT32, 1, dim, seed=4, dist="grid")[0]
    ri, pi = R.batch_iterator(q), P.batch_iterator(q)
    seen_r, seen_p = [], []
    while ri.has_next():
        assert pi.has_next()
        rl, rs, _ = ri.next(37)
        pl, ps, _ = pi.next(37)
        assert np.array_equal(rs, ps)          # score sequence identical
        seen_r += rl.tolist()
        seen_p += pl.tolist()
        # label sets may differ only inside the boundary tie group: compare via scores of labels
    assert not pi.has_next()
    assert sorted(seen_r) == sorted(seen_p) == list(range(n))
    ri.close()
    pi.close()


def test_timeout_and_range_errors(ref, port):
    X, R, P = _flat_pair(ref, port, FLOAT32, 4, L2, 50, seed=1)
    q = X[0]
    ref.set_timeout(1)
    try:
        rl, rs, rc = R.topk(q, 5)
        rrl, _, rrc = R.range(q, 1.0)
    finally:
        ref.set_timeout(0)
    pl, ps, pc = P.topk(q, 5, timeout=1)
    prl, _, prc = P.range(q, 1.0, timeout=1)
    assert rc == pc == 1 and len(rl) == len(pl) == 0
    assert rrc == prc == 1 and len(rrl) == len(prl) == 0
    with pytest.raises(RuntimeError):
        R.range(q, -1.0)
    with pytest.raises(RuntimeError):
        P.range(q, -1.0)


# ---- HNSW: the C restatement (oracle/vs_oracle_hnsw.c) against the unmodified reference ----
@pytest.mark.parametrize("metric", [0, 1, 2], ids=["L2", "IP", "Cosine"])
def test_hnsw_port_builds_the_reference_graph(port, ref, metric):
    """Same levels (the reference's std::default_random_engine stream), entry point and link lists in
    order; same top-k and range results. fp32 (the only type the reference harness exports graphs for)."""
    from datagen import make_vectors
    n, dim, M, efc = 2500, 20, 6, 40
    X = make_vectors(0, n, dim, seed=31 + metric)
    Q = make_vectors(0, 10, dim, seed=32 + metric)
    R = ref.RefIndex(0, dim, metric, algo="hnsw", M=M, ef_construction=efc, ef_runtime=10)
    R.add_many(X)
    P = port.PortHnsw(0, dim, metric, M=M, ef_construction=efc, ef_runtime=10)
    P.add_many(X)
    gr, gp = R.hnsw_export(), P.export()
    assert np.array_equal(gr["levels"], gp["levels"])
    assert (gr["entry"], gr["max_level"]) == (gp["entry"], gp["max_level"])
    assert gr["max_level"] >= 2, "fixture should exercise upper levels"
    for lvl in range(gr["max_level"] + 1):
        assert np.array_equal(gr["counts"][lvl], gp["counts"][lvl]), lvl
        assert np.array_equal(gr["links"][lvl], gp["links"][lvl]), lvl
    for ef in (0, 10, 50):
        for q in Q:
            rl, rs, _ = R.topk(q, 8, ef_runtime=ef)
            pl, ps, _ = P.topk(q, 8, ef_runtime=ef)
            assert np.array_equal(rl, pl) and np.array_equal(rs, ps)
    for q in Q:
        radius = max(float(R.topk(q, 8, ef_runtime=50)[1][5]), 0.0)
        rl, rs, _ = R.range(q, radius)
        pl, ps, _ = P.range(q, radius)
        o = np.lexsort((rl, rs))
        assert np.array_equal(rl[o], pl) and np.array_equal(rs[o], ps)
    R.close()
    P.close()


@pytest.mark.parametrize("vtype,metric,dim", [(1, 0, 12), (2, 1, 64), (2, 0, 40), (3, 2, 48), (4, 2, 64), (5, 0, 33),
                                              (0, 1, 5)], ids=lambda v: str(v))
def test_hnsw_port_queries_match_reference_all_types(port, ref, vtype, metric, dim):
    from datagen import make_vectors
    feats = ref.host_features()
    if vtype == 2 and metric != 0 and "avx512_bf16" not in feats:
        pytest.skip("host lacks avx512_bf16")
    ref.set_disabled_features("avx512_fp16")
    n = 1200
    X = make_vectors(vtype, n, dim, seed=400 + vtype)
    Q = make_vectors(vtype, 8, dim, seed=401 + vtype)
    if metric == 2:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    R = ref.RefIndex(vtype, dim, metric, algo="hnsw", M=5, ef_construction=30, ef_runtime=10)
    R.add_many(X)
    P = port.PortHnsw(vtype, dim, metric, M=5, ef_construction=30, ef_runtime=10)
    P.add_many(X)
    for ef in (10, 40):
        for q in Q:
            rl, rs, _ = R.topk(q, 6, ef_runtime=ef)
            pl, ps, _ = P.topk(q, 6, ef_runtime=ef)
            assert np.array_equal(rl, pl) and np.array_equal(rs, ps), (vtype, metric, ef)
    R.close()
    P.close()
    ref.set_disabled_features()


@pytest.mark.parametrize("metric", [0, 2], ids=["L2", "Cosine"])
def test_hnsw_port_batch_iterator_matches_reference(port, ref, metric):
    """HNSW_BatchIterator: same batches (labels and scores, in order) for several batch-size schedules, incl. reset."""
    from datagen import make_vectors
    n, dim = 800, 16
    X = make_vectors(0, n, dim, seed=71 + metric)
    Q = make_vectors(0, 4, dim, seed=72 + metric)
    R = ref.RefIndex(0, dim, metric, algo="hnsw", M=6, ef_construction=40, ef_runtime=10)
    R.add_many(X)
    P = port.PortHnsw(0, dim, metric, M=6, ef_construction=40, ef_runtime=10)
    P.add_many(X)
    for q in Q:
        for sched in ([5, 5, 5, 20, 1, 100], [1, 2, 3], [50, 50], [1000]):
            ri, pi = R.batch_iterator(q), P.batch_iterator(q)
            for rounds in range(2):
                for nres in sched:
                    assert ri.has_next() == pi.has_next()
                    rl, rs, _ = ri.next(nres)
                    pl, ps, _ = pi.next(nres)
                    assert np.array_equal(rl, pl) and np.array_equal(rs, ps), (sched, nres)
                assert ri.has_next() == pi.has_next()
                ri.reset()
                pi.reset()
            ri.close()
            pi.close()
    R.close()
    P.close()


@pytest.mark.parametrize("vtype,metric", [(0, 0), (0, 1), (0, 2), (2, 0), (3, 1), (4, 2), (5, 0), (1, 0)], ids=lambda v: str(v))
def test_hnsw_multi_port_matches_reference(port, ref, vtype, metric):
    """HNSWIndex_Multi (hnsw_multi.h): eight vectors per label; the restated label-keyed result set
    (updatable_max_heap semantics) and the unique range container give the unmodified reference's top-k (several k / ef,
    k up to the label count) and range replies — labels, order and scores. First half of SURVEY §8 row f2 for HNSW: the
    checker exists and is pinned before the device kernel is written."""
    from datagen import make_vectors
    feats = ref.host_features()
    if vtype == 2 and metric != 0 and "avx512_bf16" not in feats:
        pytest.skip("host lacks avx512_bf16")
    ref.set_disabled_features("avx512_fp16")
    n, dim = 1200, 16
    labels = (np.arange(n) % 150).astype(np.uint64)
    X = make_vectors(vtype, n, dim, seed=3)
    Q = make_vectors(vtype, 10, dim, seed=4)
    if metric == 2:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    R = ref.RefIndex(vtype, dim, metric, multi=True, algo="hnsw", M=8, ef_construction=40, ef_runtime=20)
    R.add_many(X, labels=labels)
    P = port.PortHnsw(vtype, dim, metric, M=8, ef_construction=40, ef_runtime=20, multi=True)
    P.add_many(X, labels=labels)
    for q in Q:
        for k, ef in ((10, 0), (5, 50), (40, 0), (150, 0)):
            rl, rs, _ = R.topk(q, k, ef_runtime=ef)
            pl, ps, _ = P.topk(q, k, ef_runtime=ef)
            assert len(set(rl.tolist())) == len(rl)                       # each label once
            assert np.array_equal(rl, pl) and np.array_equal(rs, ps), (vtype, metric, k, ef)
        rs20 = R.topk(q, 20)[1]
        radius = float(rs20[-1]) if rs20[-1] > 0 else 0.3
        rl, rs, _ = R.range(q, radius)
        pl, ps, _ = P.range(q, radius)
        o = np.lexsort((rl, rs))
        assert np.array_equal(rl[o], pl) and np.array_equal(rs[o], ps), (vtype, metric, "range")
    R.close()
    P.close()
    ref.set_disabled_features()


@pytest.mark.parametrize("metric", [0, 2], ids=["L2", "Cosine"])
def test_hnsw_multi_port_batch_iterator_matches_reference(port, ref, metric):
    """HNSWMulti_BatchIterator (hnsw_multi_batch_iterator.h:39-99): label-keyed result set, labels already handed out are
    skipped when refilling from the extras or admitting from the graph; same batches as the unmodified reference for
    several batch-size schedules, incl. reset, until every label was returned exactly once."""
    from datagen import make_vectors
    n, dim = 900, 16
    labels = (np.arange(n) % 100).astype(np.uint64)          # nine vectors per label
    X = make_vectors(0, n, dim, seed=81 + metric)
    Q = make_vectors(0, 4, dim, seed=82 + metric)
    R = ref.RefIndex(0, dim, metric, multi=True, algo="hnsw", M=6, ef_construction=40, ef_runtime=10)
    R.add_many(X, labels=labels)
    P = port.PortHnsw(0, dim, metric, M=6, ef_construction=40, ef_runtime=10, multi=True)
    P.add_many(X, labels=labels)
    for q in Q:
        for sched in ([5, 5, 5, 20, 1, 100], [1, 2, 3], [50, 50], [1000]):
            ri, pi = R.batch_iterator(q), P.batch_iterator(q)
            for rounds in range(2):
                seen = []
                for nres in sched:
                    assert ri.has_next() == pi.has_next()
                    rl, rs, _ = ri.next(nres)
                    pl, ps, _ = pi.next(nres)
                    assert np.array_equal(rl, pl) and np.array_equal(rs, ps), (sched, nres)
                    seen += pl.tolist()
                assert len(seen) == len(set(seen))
                assert ri.has_next() == pi.has_next()
                ri.reset()
                pi.reset()
            ri.close()
            pi.close()
    R.close()
    P.close()


def test_reference_loader_and_writer_agree_with_the_restated_file_reader(port, ref, tmp_path):
    """The reference's own HNSW file loader / writer (BUILD_TESTS variant of oracle/_ref): it loads its V3 fixture with a
    valid integrity check, the restated reader counts the same unidirectional edges, and the V4 file the reference then
    writes parses back to the same index — so oracle/port.py::read_hnsw_file is pinned on both encodings."""
    if not ref.bt_available():
        pytest.skip("oracle/_ref/libvecsim_ref_bt.so not built (make -C oracle ref_bt)")
    fixture = os.path.join(HERE, "golden", "ref_hnsw_1k_d4_single.v3")
    R = ref.RefFileIndex(fixture)
    ok, double_conn, unidir = R.integrity()
    f3 = port.read_hnsw_file(fixture)
    assert ok == 1 and R.size() == 1001 and unidir == f3["incoming"]
    assert double_conn == sum(int(c.sum()) for c in f3["counts"]) - unidir   # counted once per direction
    out = str(tmp_path / "ref_written.hnsw_v4")
    R.save(out)
    f4 = port.read_hnsw_file(out)
    assert f4["version"] == 4 and f4["n"] == 1001 and f4["incoming"] == f3["incoming"]
    for key in ("dim", "type", "metric", "M", "M0", "ef_construction", "ef_runtime", "epsilon", "block_size", "entry",
                "max_level", "num_deleted"):
        assert f4[key] == f3[key], key
    assert np.array_equal(f4["labels"], f3["labels"]) and np.array_equal(f4["vectors"], f3["vectors"])
    assert np.array_equal(f4["levels"], f3["levels"])
    for lvl in range(len(f3["links"])):
        assert np.array_equal(f4["counts"][lvl], f3["counts"][lvl]) and np.array_equal(f4["links"][lvl], f3["links"][lvl])
    # the loaded index answers like the golden case
    gold = np.load(os.path.join(HERE, "golden", "hnsw_file_case.npz"))
    for i, q in enumerate(gold["Q"]):
        l, s, _ = R.topk(q, 10, ef_runtime=50)
        assert np.array_equal(l.astype(np.int64), gold["labels_ef50"][i]) and np.array_equal(s, gold["scores_ef50"][i])
    R.close()
