"""BASELINE.json's configurations at FULL size on one B200, checked through properties that do not need the
oracle to scan the whole store (it would take hours on the CPU):

  * planted rows — copies of the queries (and scaled copies) hidden at known ids must come back first with
    the score the oracle computes for that single pair;
  * the tensor-core path and the exact scan are two independent implementations: identical ids, order and
    scores on the same queries;
  * every returned (id, score) is re-scored by the oracle from the row read back from the device
    (a spot check of 1 000s of pairs), lists are sorted ascending (score, id), ids are unique;
  * a sharded run (rows split in two stores, per-shard top-K merged) equals the single-store run.

configs[0] (100 k x 128, single query) is small enough for the oracle to scan outright.
Skipped when the device has less free memory than the configuration needs."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from datagen import from_bf16, to_bf16

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    yield c
    c.set_topk_mode(0)


def _free_gb():
    free, _ = torch.cuda.mem_get_info()
    return free / 1e9


def _ingest(capi, tname, vtype, metric, n, dim, planted):
    """rows generated on the device in chunks (the bench's generator); `planted`: {id: row as numpy}."""
    import bench as sharded  # gen_chunk_torch / gen_queries_numpy live in bench.py
    G = capi.BFIndex(capi.BFParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=n, blockSize=1024))
    dev = torch.device("cuda", 0)
    chunk = 500_000
    for c0 in range(0, n, chunk):
        rows = min(chunk, n - c0)
        x = sharded.gen_chunk_torch(torch, tname, c0 // chunk, chunk, dim, dev)[:rows].contiguous()
        for pid, row in planted.items():
            if c0 <= pid < c0 + rows:
                t = torch.from_numpy(row.view(np.uint8).copy()).to(dev)
                x.view(torch.uint8).reshape(rows, -1)[pid - c0].copy_(t)
        torch.cuda.synchronize()
        G.add_device_rows(x.data_ptr(), x.stride(0) * x.element_size(), rows, c0)
        del x
    torch.cuda.synchronize()
    assert G.index_size() == n
    return G


def _read_rows(capi, G, ids, row_bytes):
    L = C.CDLL(os.path.join(ROOT, "vectorsimilarity_b200", "libvsgpu.so"))
    L.vsgpu_store_read.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    out = np.zeros((len(ids), row_bytes), dtype=np.uint8)
    for i, rid in enumerate(ids):
        assert L.vsgpu_store_read(G.device_store(), int(rid), 1, out[i].ctypes.data, row_bytes, None) == 0
    return out


def _check_lists(labels, scores):
    for i in range(labels.shape[0]):
        assert len(set(labels[i])) == labels.shape[1], "duplicate ids"
        key = list(zip(scores[i], labels[i]))
        assert key == sorted(key), "not ascending (score, id)"


def test_config0_flat_fp32_l2_100k_single_query_vs_oracle(capi, port):
    n, dim, k = 100_000, 128, 10
    rng = np.random.default_rng(47)
    X = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    G = capi.BFIndex(capi.BFParams(type=0, dim=dim, metric=0, multi=False, initialCapacity=n, blockSize=1024))
    G.add_vectors(X)
    P = port.PortIndex(0, dim, 0)
    P.add_many(X)
    for s in range(5):
        q = np.random.default_rng(48 + s).uniform(-1, 1, dim).astype(np.float32)
        gl, gs = G.knn_query(q, k)
        pl, ps, _ = P.topk(q, k)
        assert np.array_equal(gl[0], pl.astype(np.int64)) and np.array_equal(gs[0], ps)
    G.close()
    P.close()


def test_config1_flat_fp32_ip_10M_d768_k100(capi, port):
    n, dim, k, nq = 10_000_000, 768, 100, 64
    if _free_gb() < 60:
        pytest.skip("needs ~50 GB of HBM")
    import bench as sharded
    Q = sharded.gen_queries_numpy("fp32", nq, dim)
    planted = {1234567 + 31 * i: Q[i].copy() for i in range(8)}          # exact copies: score 1 - |q|^2
    planted.update({9_000_001 + 7 * i: (Q[i] * np.float32(0.5)).astype(np.float32) for i in range(8)})
    G = _ingest(capi, "fp32", 0, 1, n, dim, planted)
    capi.set_topk_mode(2)
    tl, ts = G.knn_batch(Q, k)
    st = G.last_query_stats()
    assert st["path"] == 1
    capi.set_topk_mode(1)
    el, es = G.knn_batch(Q[:16], k)
    assert np.array_equal(tl[:16], el) and np.array_equal(ts[:16], es), "tensor path != exact scan"
    _check_lists(tl, ts)
    for i in range(8):
        assert tl[i][0] == 1234567 + 31 * i                              # the planted copy wins
        assert ts[i][0] == port.distance(0, 1, Q[i], Q[i])
        assert 9_000_001 + 7 * i in tl[i][:3]
    # re-score a sample of returned pairs with the oracle from the rows as stored on the device
    sample = [(i, j) for i in range(0, nq, 4) for j in (0, 1, 50, 99)]
    rows = _read_rows(capi, G, [tl[i][j] for i, j in sample], dim * 4).view(np.float32)
    for (i, j), row in zip(sample, rows):
        assert ts[i][j] == port.distance(0, 1, row, Q[i])
    G.close()


def test_config3_flat_bf16_ip_20M_d1024_k100(capi, port):
    n, dim, k, nq = 20_000_000, 1024, 100, 64
    if _free_gb() < 60:
        pytest.skip("needs ~45 GB of HBM")
    import bench as sharded
    port.set_tier(port.TIER_AVX512)
    Q = sharded.gen_queries_numpy("bf16", nq, dim)
    planted = {7_654_321 + 13 * i: Q[i].copy() for i in range(8)}
    G = _ingest(capi, "bf16", 2, 1, n, dim, planted)
    capi.set_topk_mode(2)
    tl, ts = G.knn_batch(Q, k)
    assert G.last_query_stats()["path"] == 1
    capi.set_topk_mode(1)
    el, es = G.knn_batch(Q[:8], k)
    assert np.array_equal(tl[:8], el) and np.array_equal(ts[:8], es), "tensor path != exact scan"
    _check_lists(tl, ts)
    for i in range(8):
        assert tl[i][0] == 7_654_321 + 13 * i
        assert ts[i][0] == port.distance(2, 1, Q[i], Q[i])
    sample = [(i, j) for i in range(0, nq, 8) for j in (0, 1, 99)]
    rows = _read_rows(capi, G, [tl[i][j] for i, j in sample], dim * 2).view(np.uint16)
    for (i, j), row in zip(sample, rows):
        assert ts[i][j] == port.distance(2, 1, row, Q[i])
    G.close()


def test_config2_flat_int8_cosine_shard_6p25M_d512_k10(capi, port):
    """configs[2] is 50 M rows over 8 GPUs: one GPU's shard (6.25 M rows), plus the two-store merge."""
    n, dim, k, nq = 6_250_000, 512, 10, 32
    if _free_gb() < 12:
        pytest.skip("needs ~8 GB of HBM")
    import bench as sharded
    Q = sharded.gen_queries_numpy("int8", nq, dim)
    planted = {3_000_003 + 11 * i: Q[i].copy() for i in range(4)}
    G = _ingest(capi, "int8", 4, 2, n, dim, planted)
    gl, gs = G.knn_batch(Q, k)
    _check_lists(gl, gs)
    for i in range(4):
        assert gl[i][0] == 3_000_003 + 11 * i
    sample = [(i, j) for i in range(0, nq, 4) for j in (0, 1, 9)]
    rows = _read_rows(capi, G, [gl[i][j] for i, j in sample], dim + 4)
    for (i, j), row in zip(sample, rows):
        qn = np.zeros(dim + 4, dtype=np.uint8)
        qn[:dim] = Q[i].view(np.uint8)
        port.normalize(4, dim, qn)                                       # appends the query's norm
        assert gs[i][j] == port.distance(4, 2, row, qn, dim=dim)
    # shard merge: the same rows split over two stores give the same global top-K
    half = n // 2
    H = [capi.BFIndex(capi.BFParams(type=4, dim=dim, metric=2, multi=False, initialCapacity=half, blockSize=1024))
         for _ in range(2)]
    dev = torch.device("cuda", 0)
    chunk = 500_000
    for c0 in range(0, n, chunk):
        rows_n = min(chunk, n - c0)
        x = sharded.gen_chunk_torch(torch, "int8", c0 // chunk, chunk, dim, dev)[:rows_n].contiguous()
        for pid, row in planted.items():
            if c0 <= pid < c0 + rows_n:
                x[pid - c0].copy_(torch.from_numpy(row.copy()).to(dev))
        torch.cuda.synchronize()
        for s, (lo, hi) in enumerate(((0, half), (half, n))):
            a, b = max(lo, c0), min(hi, c0 + rows_n)
            if a < b:
                part = x[a - c0:b - c0].contiguous()
                torch.cuda.synchronize()
                H[s].add_device_rows(part.data_ptr(), part.stride(0), b - a, a)
    parts = [h.knn_batch(Q, k) for h in H]
    S = np.stack([p[1] for p in parts]).astype(np.float32)
    Lb = np.stack([p[0].view(np.uint64) for p in parts])
    from vectorsimilarity_b200.sharded import merge_topk_host
    ms, ml = merge_topk_host(S, Lb, k)
    assert np.array_equal(ml.view(np.int64), gl) and np.array_equal(ms.astype(np.float64), gs)
    for h in H:
        h.close()
    G.close()


def test_config4_hnsw_fp32_l2_1M_reference_graph(capi):
    """configs[4]: the graph the unmodified reference built over 1 M vectors (hnsw_cache/, made by
    scripts/make_hnsw_cfg5.py; git-ignored, shipped with the snapshot), searched on the device: every id and score equals
    the reference's recorded answer. Without the file the same check runs on a 60 k-node graph that the unmodified
    reference (oracle/_ref) builds on this box — the test never skips."""
    path = os.path.join(ROOT, "hnsw_cache", "cfg5_graph_1000000.npz")
    dim, M = 128, 16
    rng = np.random.default_rng(47)
    if os.path.exists(path):
        g = np.load(path)
        n = 1_000_000
        X = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
        Q = rng.uniform(-1, 1, (256, dim)).astype(np.float32)
        levels, l0, upper = (np.ascontiguousarray(g[k]) for k in ("levels", "l0", "upper"))
        entry, max_level = int(g["entry"][0]), int(g["entry"][1])
        want_l, want_s = g["ref_labels"], g["ref_scores"]
    else:
        from oracle import ref
        n = 60_000
        X = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
        Q = rng.uniform(-1, 1, (256, dim)).astype(np.float32)
        ref.lib()
        R = ref.RefIndex(0, dim, 0, algo="hnsw", M=M, ef_construction=200, ef_runtime=64)
        R.add_many(X)
        e = R.hnsw_export()
        levels = np.ascontiguousarray(e["levels"], dtype=np.uint32)
        l0 = np.zeros((n, 2 * M + 1), dtype=np.uint32)
        l0[:, 0] = e["counts"][0]
        l0[:, 1:] = np.where(np.arange(2 * M)[None, :] < e["counts"][0][:, None], e["links"][0], 0)
        recs = []
        for i in np.nonzero(levels)[0]:
            for lvl in range(1, int(levels[i]) + 1):
                r = np.zeros(M + 1, dtype=np.uint32)
                c = int(e["counts"][lvl][i])
                r[0] = c
                r[1:1 + c] = e["links"][lvl][i][:c]
                recs.append(r)
        upper = np.stack(recs) if recs else np.zeros((0, M + 1), dtype=np.uint32)
        entry, max_level = int(e["entry"]), int(e["max_level"])
        want_l = np.zeros((256, 10), dtype=np.int64)
        want_s = np.zeros((256, 10), dtype=np.float64)
        for i in range(256):
            want_l[i], want_s[i], _ = R.topk(Q[i], 10, ef_runtime=64)
        R.close()
    G = capi.HNSWIndex(capi.HNSWParams(type=0, dim=dim, metric=0, multi=False, initialCapacity=n, blockSize=1024, M=M,
                                       efConstruction=200, efRuntime=64, epsilon=0.01))
    rc = capi.lib().VecSimGPU_HNSWImportGraph(G._h, X.ctypes.data, 1, n, None, levels.ctypes.data, l0.ctypes.data,
                                              upper.ctypes.data if len(upper) else None, len(upper), entry, max_level)
    assert rc == 0
    labels, scores = G.knn_batch(Q, 10)
    assert np.array_equal(labels, want_l)
    assert np.array_equal(scores, want_s)
    G.close()
