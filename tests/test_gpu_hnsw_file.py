"""Serialized HNSW index files through the C-ABI: the reference's own V3 fixture is loaded into the
device store + device graph and answers exactly as the unmodified reference does on that graph (golden case from
tests/golden/make_hnsw_file_golden.py); the V4 writer's output is parsed back by the oracle's restated reader and
reloaded. Mirrors tests/unit/test_hnsw.cpp:1915-2059 (HNSWSerialization*, HNSWSerializationV3)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "ref_hnsw_1k_d4_single.v3")
GOLD = np.load(os.path.join(HERE, "golden", "hnsw_file_case.npz"))


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


def file_records(f):
    n, M = f["n"], f["M"]
    l0 = np.zeros((n, 2 * M + 1), dtype=np.uint32)
    l0[:, 0] = f["counts"][0]
    l0[:, 1:] = np.where(np.arange(2 * M)[None, :] < f["counts"][0][:, None], f["links"][0], 0)
    recs = []
    for i in np.nonzero(f["levels"])[0]:
        for lvl in range(1, int(f["levels"][i]) + 1):
            r = np.zeros(M + 1, dtype=np.uint32)
            c = int(f["counts"][lvl][i])
            r[0] = c
            r[1:1 + c] = f["links"][lvl][i][:c]
            recs.append(r)
    return l0, (np.stack(recs) if recs else np.zeros((0, M + 1), dtype=np.uint32))


def masked(rec):
    rec = rec.copy()
    w = rec.shape[1] - 1
    rec[:, 1:] = np.where(np.arange(w)[None, :] < rec[:, :1], rec[:, 1:], 0)
    return rec


def test_reference_v3_fixture_loads_and_answers_like_the_reference(capi, port):
    G = capi.HNSWIndex.load(FIXTURE)
    info = dict(G.debug_info())
    # what tests/unit/test_hnsw.cpp:2030-2045 asserts after loading this file
    assert info["ALGORITHM"] == "HNSW" and info["M"] == 8 and info["IS_MULTI_VALUE"] == 0 and info["BLOCK_SIZE"] == 2
    assert info["EF_CONSTRUCTION"] == 10 and info["EF_RUNTIME"] == 10 and info["INDEX_SIZE"] == 1001
    assert info["METRIC"] == "L2" and info["TYPE"] == "FLOAT32" and info["DIMENSION"] == 4
    assert info["INDEX_LABEL_COUNT"] == 1001 and info["EPSILON"] == 0.004
    # the device graph is the file's graph
    f = port.read_hnsw_file(FIXTURE)
    g = G.export_graph(1001)
    l0, upper = file_records(f)
    assert g["entry"] == f["entry"] and g["max_level"] == f["max_level"] and np.array_equal(g["levels"], f["levels"])
    assert np.array_equal(masked(g["l0"]), l0) and np.array_equal(masked(g["upper"]), upper)
    # and it answers as the unmodified reference does on that graph
    for ef in (10, 50):
        G.set_ef(ef)
        labels, scores = G.knn_batch(GOLD["Q"], 10)
        assert np.array_equal(labels, GOLD[f"labels_ef{ef}"]) and np.array_equal(scores, GOLD[f"scores_ef{ef}"])
    l, s = G.range_query(GOLD["Q"][0], 0.05)
    assert np.array_equal(l[0], GOLD["range_labels"]) and np.array_equal(s[0], GOLD["range_scores"])
    # "check the functionality of the loaded index": add and delete (test_hnsw.cpp:2049-2056)
    assert G.add_vector(np.full(4, 1001, dtype=np.float32), 1001) == 1
    assert G.delete_vector(1) == 1
    assert G.index_size() == 1001
    l, _ = G.knn_query(np.full(4, 1001, dtype=np.float32), 1)
    assert l[0][0] == 1001
    G.close()


@pytest.mark.parametrize("vtype,metric", [(0, 0), (0, 2), (4, 2), (2, 1)])
def test_save_then_parse_and_reload(capi, port, tmp_path, vtype, metric):
    from datagen import make_vectors
    n, dim, M = 700, 16, 6
    X = make_vectors(vtype, n, dim, seed=31)
    Q = make_vectors(vtype, 8, dim, seed=32)
    labels = (np.arange(n) * 3 + 5).astype(np.uint64)
    G = capi.HNSWIndex(capi.HNSWParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=0, blockSize=64, M=M,
                                       efConstruction=30, efRuntime=15, epsilon=0.02))
    G.add_vectors(X, labels=labels)
    for lab in (5, 8, 305):
        assert G.delete_vector(lab) == 1
    want_l, want_s = G.knn_batch(Q, 10)
    path = str(tmp_path / "index.hnsw_v4")
    G.save_index(path)
    # the oracle's reader (restating the reference's restore code) accepts the file and finds the same index in it
    f = port.read_hnsw_file(path)
    assert (f["version"], f["n"], f["dim"], f["type"], f["metric"], f["M"], f["M0"]) == (4, n, dim, vtype, metric, M, 2 * M)
    assert (f["ef_construction"], f["ef_runtime"], f["epsilon"], f["block_size"], f["num_deleted"]) == (30, 15, 0.02, 64, 3)
    assert f["capacity"] % 64 == 0 and f["capacity"] >= n and abs(f["mult"] - 1 / np.log(M)) < 1e-15
    assert np.array_equal(f["labels"], labels)
    assert np.nonzero(f["flags"])[0].tolist() == [0, 1, 100] and set(f["flags"].tolist()) == {0, 1}
    g = G.export_graph(n)
    l0, upper = file_records(f)
    assert g["entry"] == f["entry"] and g["max_level"] == f["max_level"] and np.array_equal(g["levels"], f["levels"])
    assert np.array_equal(masked(g["l0"]), l0) and np.array_equal(masked(g["upper"]), upper)
    # stored rows are the processed rows (normalised for Cosine; int8 Cosine carries its norm)
    P = port.PortHnsw(vtype, dim, metric, M=M, ef_construction=30, ef_runtime=15)
    P.add_many(X, labels=labels)
    assert f["vectors"].shape[1] == port.stored_size(vtype, metric, dim)
    # incoming lists = edges that are not mutual
    links0 = [set(f["links"][0][i, :f["counts"][0][i]].tolist()) for i in range(n)]
    one_way = sum(1 for u in range(n) for v in links0[u] if u not in links0[v])
    one_way_upper = 0
    for lvl in range(1, len(f["links"])):
        ls = [set(f["links"][lvl][i, :f["counts"][lvl][i]].tolist()) for i in range(n)]
        one_way_upper += sum(1 for u in range(n) if f["levels"][u] >= lvl for v in ls[u] if u not in ls[v])
    assert f["incoming"] == one_way + one_way_upper
    # reload: same answers, deleted labels stay deleted
    G2 = capi.HNSWIndex.load(path)
    assert G2.index_size() == n - 3
    got_l, got_s = G2.knn_batch(Q, 10)
    assert np.array_equal(got_l, want_l) and np.array_equal(got_s, want_s)
    assert G2.delete_vector(5) == 0 and G2.delete_vector(11) == 1
    P.close(); G.close(); G2.close()


def test_bad_files_are_refused(capi, tmp_path):
    L = capi.lib()
    assert not L.VecSimGPU_HNSWLoadIndex(os.fsencode(str(tmp_path / "missing")))
    assert b"Cannot open file" in L.VecSimGPU_LastError()
    raw = open(FIXTURE, "rb").read()
    for name, data, msg in (("trunc", raw[:30000], b"truncated"), ("old", b"\x02\x00\x00\x00" + raw[4:], b"deprecated"),
                            ("new", b"\x09\x00\x00\x00" + raw[4:], b"bad encoding"), ("flat", raw[:4] + b"\x00\x00\x00\x00" + raw[8:], b"Expected HNSW")):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        assert not L.VecSimGPU_HNSWLoadIndex(os.fsencode(p)), name
        assert msg in L.VecSimGPU_LastError(), (name, L.VecSimGPU_LastError())


@pytest.mark.parametrize("metric", [0, 2])
def test_reference_loads_what_we_write(capi, tmp_path, metric):
    """The file VecSimGPU_HNSWSaveIndex writes is loaded by the REFERENCE's own loader (BUILD_TESTS variant of
    oracle/_ref, HNSWFactory::NewIndex(location)): its integrity check passes, it holds the same number of vectors, and
    it answers top-k exactly as the device index that wrote the file — deleted marks included."""
    from oracle import ref
    if not ref.bt_available():
        pytest.skip("oracle/_ref/libvecsim_ref_bt.so not built")
    from datagen import make_vectors
    n, dim, M = 900, 16, 6
    X = make_vectors(0, n, dim, seed=41)
    Q = make_vectors(0, 12, dim, seed=42)
    G = capi.HNSWIndex(capi.HNSWParams(type=0, dim=dim, metric=metric, multi=False, initialCapacity=0, blockSize=64, M=M,
                                       efConstruction=30, efRuntime=15, epsilon=0.02))
    G.add_vectors(X, labels=(np.arange(n) + 100).astype(np.uint64))
    for lab in (100, 350, 999):
        assert G.delete_vector(lab) == 1
    path = str(tmp_path / "ours.hnsw_v4")
    G.save_index(path)
    R = ref.RefFileIndex(path)
    ok, double_conn, unidir = R.integrity()
    assert ok == 1, "the reference's checkIntegrity rejects the file"
    assert R.size() == n            # the reference counts mark-deleted nodes in indexSize (hnsw.h:367-369)
    for ef in (15, 60):
        G.set_ef(ef)
        labels, scores = G.knn_batch(Q, 10)
        for i, q in enumerate(Q):
            rl, rs, _ = R.topk(q, 10, ef_runtime=ef)
            assert np.array_equal(labels[i], rl.astype(np.int64)) and np.array_equal(scores[i], rs), (metric, ef, i)
    R.close()
    G.close()


def test_overwritten_label_stays_overwritten_after_reload(capi, tmp_path):
    """an upsert tombstones the old node and appends a new one with the same label; the saved file holds
    both (old flagged DELETE_MARK). After loading, the stale vector must stay dead on the device — one result per label,
    the new vector's score — and the deleted count must match what was saved."""
    from datagen import make_vectors
    n, dim = 400, 8
    X = make_vectors(0, n, dim, seed=7)
    G = capi.HNSWIndex(capi.HNSWParams(type=0, dim=dim, metric=0, multi=False, initialCapacity=0, blockSize=32, M=6,
                                       efConstruction=30, efRuntime=400, epsilon=0.01))
    G.add_vectors(X)
    far = np.full(dim, 50.0, dtype=np.float32)
    assert G.add_vector(far, 17) == 0           # overwrite: label 17 now lives far away
    assert G.delete_vector(23) == 1
    assert G.add_vector(X[23], 23) == 1          # deleted, then re-added: a flagged id and a later live id share label 23
    path = str(tmp_path / "upsert.hnsw_v4")
    G.save_index(path)
    G2 = capi.HNSWIndex.load(path)
    assert G2.index_size() == G.index_size() == n
    assert dict(G2.debug_info())["NUMBER_OF_MARKED_DELETED"] == dict(G.debug_info())["NUMBER_OF_MARKED_DELETED"] == 2
    for q, lab in ((X[17], 17), (X[23], 23), (far, 17)):
        l1, s1 = G.knn_query(q, n)
        l2, s2 = G2.knn_query(q, n)
        assert np.array_equal(l1, l2) and np.array_equal(s1, s2)
        assert (l2[0] == lab).sum() == 1          # never twice
    l, s = G2.knn_query(X[17], 1)
    assert l[0][0] != 17                          # the stale copy of label 17 is not the nearest neighbour of its old self
    G.close()
    G2.close()


def test_corrupted_graph_section_is_refused(capi, tmp_path):
    """link ids, the entry point and levels are validated before anything reaches the device."""
    L = capi.lib()
    from datagen import make_vectors
    G = capi.HNSWIndex(capi.HNSWParams(type=0, dim=4, metric=0, multi=False, initialCapacity=0, blockSize=16, M=4,
                                       efConstruction=20, efRuntime=10, epsilon=0.01))
    G.add_vectors(make_vectors(0, 64, 4, seed=3))
    path = str(tmp_path / "ok.hnsw_v4")
    G.save_index(path)
    G.close()
    raw = bytearray(open(path, "rb").read())
    # header: version, algo (2 x int) | dim (size_t) | type, metric (2 x int) | blockSize (size_t) | multi (bool) |
    # initialCapacity, M, M0, efC, ef (5 x size_t) | epsilon, mult (2 x double) | n, num_deleted, max_level (3 x size_t) | entry (u32)
    off_n = 4 + 4 + 8 + 4 + 4 + 8 + 1 + 5 * 8 + 2 * 8
    off_entry = off_n + 3 * 8
    bad_entry = bytearray(raw)
    bad_entry[off_entry:off_entry + 4] = (1000).to_bytes(4, "little")
    huge_n = bytearray(raw)
    huge_n[off_n:off_n + 8] = (1 << 61).to_bytes(8, "little")
    # first link id of node 0 (V4: labels+flags, vectors, then per block: len u32, per node: level size_t, count u16, links)
    off_graph = off_entry + 4 + 64 * 9 + 64 * 16
    bad_link = bytearray(raw)
    bad_link[off_graph + 4 + 8 + 2:off_graph + 4 + 8 + 2 + 4] = (0x7fffffff).to_bytes(4, "little")
    for name, data in (("entry", bad_entry), ("count", huge_n), ("link", bad_link)):
        p = str(tmp_path / name)
        open(p, "wb").write(bytes(data))
        assert not L.VecSimGPU_HNSWLoadIndex(os.fsencode(p)), name
        assert b"corrupted" in L.VecSimGPU_LastError(), (name, L.VecSimGPU_LastError())
    G3 = capi.HNSWIndex.load(path)                 # the untouched file still loads
    assert G3.index_size() == 64
    G3.close()
