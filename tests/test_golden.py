"""The oracle port against (1) the known-answer values of the reference's own unit tests
(tests/golden/kat.json) and (2) outputs of the unmodified reference captured by
tests/golden/make_golden.py (tests/golden/flat_cases.npz). Runs anywhere (no _ref, no GPU)."""
import json
import os

import numpy as np
import pytest

from datagen import METRIC_NAMES, TYPE_NAMES, to_bf16

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat.json")))
TYPES = {n: i for i, n in enumerate(TYPE_NAMES)}
METRICS = {n: i for i, n in enumerate(METRIC_NAMES)}
NPD = {"fp32": np.float32, "fp64": np.float64, "fp16": np.float16, "int8": np.int8, "uint8": np.uint8}


def cast(tname, values):
    if tname == "bf16":
        return to_bf16(np.asarray(values, dtype=np.float32))
    return np.asarray(values, dtype=NPD[tname])


@pytest.fixture(scope="module")
def cases():
    return np.load(os.path.join(HERE, "golden", "flat_cases.npz"))


@pytest.mark.parametrize("kat", KAT["distance"], ids=lambda k: k["src"].split("/")[-1])
def test_distance_kat(port, kat):
    a, b = cast(kat["type"], kat["a"]), cast(kat["type"], kat["b"])
    for tier in (port.TIER_AVX512, port.TIER_NAIVE):
        port.set_tier(tier)
        try:
            d = port.distance(TYPES[kat["type"]], METRICS[kat["metric"]], a, b)
        finally:
            port.set_tier(port.TIER_AVX512)
        assert d == kat["expect"]


@pytest.mark.parametrize("kat", KAT["flat_topk"], ids=lambda k: k["src"].split(" ")[0].split("/")[-1])
def test_flat_topk_kat(port, kat):
    for tname in kat["types"]:
        for block_size in (1, 12, 1024):
            idx = port.PortIndex(TYPES[tname], kat["dim"], METRICS[kat["metric"]], block_size=block_size)
            for i in range(kat["n"]):
                idx.add(cast(tname, [i] * kat["dim"]), i)
            q = cast(tname, [kat["query_value"]] * kat["dim"])
            order = port.BY_ID if kat.get("order") == "BY_ID" else port.BY_SCORE
            labels, scores, code = idx.topk(q, kat["k"], order)
            assert code == 0
            if "expect_labels" in kat:
                assert labels.tolist() == kat["expect_labels"]
            if "expect_label_set" in kat:
                assert sorted(labels.tolist()) == kat["expect_label_set"]
            if "expect_scores" in kat:
                assert scores.tolist() == kat["expect_scores"]
            assert idx.topk(q, 0)[0].size == 0   # "search for nothing" (test_bruteforce.cpp:808)
            idx.close()


@pytest.mark.parametrize("vtype", range(6), ids=TYPE_NAMES)
@pytest.mark.parametrize("metric", range(3), ids=METRIC_NAMES)
def test_port_matches_captured_reference(port, cases, vtype, metric):
    if "avx512_bf16" not in str(cases["host_features"]):
        pytest.skip("fixture generated on a host without the modelled tier")
    name = f"{TYPE_NAMES[vtype]}_{METRIC_NAMES[metric]}"
    X, Q = cases[name + "_X"], cases[name + "_Q"]
    idx = port.PortIndex(vtype, X.shape[1], metric, block_size=64)
    idx.add_many(X)
    for i in range(Q.shape[0]):
        labels, scores, _ = idx.topk(Q[i], cases[name + "_labels"].shape[1])
        assert np.array_equal(labels.astype(np.int64), cases[name + "_labels"][i])
        assert np.array_equal(scores, cases[name + "_scores"][i])
        rl, _, _ = idx.range(Q[i], float(cases[name + "_radius"][i]))
        assert len(rl) == cases[name + "_range_count"][i]
    idx.close()


def test_port_matches_captured_ties(port, cases):
    idx = port.PortIndex(4, 3, 0, block_size=16)
    idx.add_many(cases["ties_X"], labels=cases["ties_lab"])
    labels, scores, _ = idx.topk(cases["ties_q"], len(cases["ties_labels"]))
    assert np.array_equal(labels, cases["ties_labels"])
    assert np.array_equal(scores, cases["ties_scores"])


# ---- HNSW: the oracle restatement against the committed reference outputs (tests/golden/hnsw_case.npz) ----
@pytest.mark.parametrize("name,metric", [("l2", 0), ("cos", 2)])
def test_hnsw_port_reproduces_golden_graph_and_results(port, name, metric):
    G = np.load(os.path.join(HERE, "golden", "hnsw_case.npz"))
    port.set_tier(port.TIER_AVX512)
    P = port.PortHnsw(0, 24, metric, M=8, ef_construction=48, ef_runtime=10)
    P.add_many(G[name + "_X"])
    g = P.export()
    assert np.array_equal(g["levels"], G[name + "_levels"])
    assert [g["entry"], g["max_level"]] == list(G[name + "_entry"])
    for lvl in range(g["max_level"] + 1):
        assert np.array_equal(g["counts"][lvl], G[f"{name}_counts{lvl}"]), lvl
        assert np.array_equal(g["links"][lvl], G[f"{name}_links{lvl}"]), lvl
    Q = G[name + "_Q"]
    for ef in (10, 40):
        for i in range(len(Q)):
            l, s, _ = P.topk(Q[i], 10, ef_runtime=ef)
            assert np.array_equal(l.astype(np.int64), G[f"{name}_labels_ef{ef}"][i])
            assert np.array_equal(s, G[f"{name}_scores_ef{ef}"][i])
    off = 0
    for i in range(len(Q)):
        n = int(G[name + "_range_n"][i])
        l, s, _ = P.range(Q[i], float(G[name + "_radius"][i]))
        assert np.array_equal(l.astype(np.int64), G[name + "_range_labels"][off:off + n])
        assert np.array_equal(s, G[name + "_range_scores"][off:off + n])
        off += n
    P.close()


def test_hnsw_file_fixture_parses_and_oracle_rebuilds_its_graph():
    """The reference's own serialized-index fixture (tests/unit/test_hnsw.cpp:1991-2059 loads it): the restated reader
    recovers the metadata that test asserts, and the oracle's builder, fed the file's vectors in id order, reproduces
    the graph stored in the file link for link — a graph written by the reference on its maintainers' machine."""
    from oracle import port
    port.build()
    f = port.read_hnsw_file(os.path.join(HERE, "golden", "ref_hnsw_1k_d4_single.v3"))
    assert (f["version"], f["n"], f["dim"], f["type"], f["metric"]) == (3, 1001, 4, 0, 0)
    assert (f["M"], f["M0"], f["ef_construction"], f["ef_runtime"], f["epsilon"], f["block_size"]) == (8, 16, 10, 10, 0.004, 2)
    assert not f["multi"] and f["num_deleted"] == 0 and np.array_equal(f["labels"], np.arange(1001))
    X = np.ascontiguousarray(f["vectors"]).view(np.float32)
    H = port.PortHnsw(0, 4, 0, M=8, ef_construction=10, ef_runtime=10)
    H.add_many(X, labels=f["labels"])
    g = H.export()
    assert g["entry"] == f["entry"] and g["max_level"] == f["max_level"] and np.array_equal(g["levels"], f["levels"])
    for lvl in range(len(f["links"])):
        assert np.array_equal(g["counts"][lvl], f["counts"][lvl])
        mask = np.arange(f["links"][lvl].shape[1])[None, :] < f["counts"][lvl][:, None]
        assert np.array_equal(np.where(mask, g["links"][lvl], 0), np.where(mask, f["links"][lvl], 0))
    # and its answers are the unmodified reference's (tests/golden/make_hnsw_file_golden.py)
    gold = np.load(os.path.join(HERE, "golden", "hnsw_file_case.npz"))
    for ef in (10, 50):
        for i, q in enumerate(gold["Q"]):
            l, s, _ = H.topk(q, 10, ef_runtime=ef)
            assert np.array_equal(l.astype(np.int64), gold[f"labels_ef{ef}"][i]) and np.array_equal(s, gold[f"scores_ef{ef}"][i])
    H.close()
