"""The sharded flat index INSIDE the C library (VecSimGPU_Configure + VecSimIndex_New(VecSimAlgo_BF), SURVEY §5 / §8b
"Additions"): one process, several device stores, per-shard top-k gathered by peer copies and merged on the root device.
The same cases as tests/nccl_parity_check.py (the torch.distributed front-end), through libvecsim_b200.so alone: results
must equal the oracle's over the whole, unsharded row set — ids, order and scores. On a one-GPU box the shards share the
device (a device may be listed several times); with more GPUs visible each shard gets its own."""
import numpy as np
import pytest

from datagen import make_vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    yield c
    c.set_device(0)
    c.set_topk_mode(0)


def _devices(capi, shards):
    have = capi.device_count()
    return [i % have for i in range(shards)]


@pytest.mark.parametrize("vtype,metric,n,dim,k,nq,mode,shards", [
    (0, 1, 20011, 96, 25, 40, 0, 3),
    (0, 0, 5003, 128, 10, 3, 1, 2),
    (4, 0, 3001, 6, 20, 9, 1, 4),       # coarse int8: many exact ties across the shards
    (2, 1, 30011, 64, 50, 64, 0, 2),
    (3, 2, 4001, 24, 7, 5, 1, 8),
    (0, 1, 140000, 64, 10, 32, 2, 3),   # every shard large enough for the tensor path (forced)
    (4, 2, 120000, 64, 10, 64, 2, 2),   # int8 cosine on the kind::i8 GEMM, per shard
])
def test_sharded_index_equals_unsharded_oracle(capi, port, vtype, metric, n, dim, k, nq, mode, shards):
    port.set_tier(port.TIER_AVX512)
    X = make_vectors(vtype, n, dim, seed=11 + vtype)
    Q = make_vectors(vtype, nq, dim, seed=12 + vtype)
    if vtype == 4 and dim < 16:
        X, Q = (X // 32).astype(np.int8), (Q // 32).astype(np.int8)
    if metric == 2 and vtype >= 4:
        X[(X == 0).all(1), 0] = 1
        Q[(Q == 0).all(1), 0] = 1
    capi.configure_devices(_devices(capi, shards))
    capi.set_topk_mode(mode)
    try:
        G = capi.BFIndex(capi.BFParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=n, blockSize=1024))
        assert G.shard_count() == shards
        assert G.add_vectors(X) == n                      # labels 0..n-1, routed by label mod shards
        assert G.index_size() == n
        P = port.PortIndex(vtype, dim, metric)
        P.add_many(X)
        labels, scores = G.knn_batch(Q, k)
        for i in range(nq):
            pl, ps, _ = P.topk(Q[i], k)
            assert np.array_equal(labels[i], pl.astype(np.int64)), (i, labels[i], pl)
            assert np.array_equal(scores[i], ps), i
        if mode == 2:
            assert G.last_query_stats()["path"] == 1
        # single query through the unchanged VecSimIndex_TopKQuery (one query never takes the tensor path)
        capi.set_topk_mode(0)
        l, s = G.knn_query(Q[0], k)
        pl, ps, _ = P.topk(Q[0], k)
        assert np.array_equal(l[0], pl.astype(np.int64)) and np.array_equal(s[0], ps)
        # range query and batch iterator across the shards
        pl, ps, _ = P.topk(Q[0], min(60, n))
        radius = float(ps[-1]) if ps[-1] >= 0 else 0.25
        for order in (capi.BY_SCORE, capi.BY_ID):
            wl, ws, _ = P.range(Q[0], radius, order=order)
            gl, gs = G.range_query(Q[0], radius, order=order)
            assert np.array_equal(gl[0], wl.astype(np.int64)) and np.array_equal(gs[0], ws), ("range", order)
        it, pit = G.create_batch_iterator(Q[0]), P.batch_iterator(Q[0])
        for bs in (1, 10, 37, 100):
            gl, gs = it.get_next_results(bs)
            wl, ws, _ = pit.next(bs)
            assert np.array_equal(gl[0], wl.astype(np.int64)) and np.array_equal(gs[0], ws), ("iterator", bs)
        it.close()
        pit.close()
        # delete / overwrite are routed to the owning shard
        for lab in (0, 1, shards, n - 1):
            assert G.delete_vector(lab) == P.delete(lab) == 1
        assert G.delete_vector(n + 5) == 0
        assert G.add_vector(X[3], 7) == 0 and P.add(X[3], 7) == 0          # overwrite label 7
        assert G.index_size() == P.size() == n - 4
        # (the tie-heavy int8 case is left out here: after deletes the reference's scan order depends on which row its
        # single store moved into each hole, a shard moves its own last row — ties AT the k-th score may differ)
        if not (vtype == 4 and dim < 16):
            labels, scores = G.knn_batch(Q[:8], k)
            for i in range(min(8, nq)):
                pl, ps, _ = P.topk(Q[i], k)
                assert np.array_equal(labels[i], pl.astype(np.int64)) and np.array_equal(scores[i], ps), ("after delete", i)
        assert abs(G.get_distance_from(9, Q[0]) - P.distance_from(9, Q[0])) == 0
        G.close()
        P.close()
    finally:
        capi.set_device(0)
        capi.set_topk_mode(0)


def test_sharded_bulk_device_ingest(capi, port):
    """VecSimGPU_AppendDeviceRows on a sharded index: rows stay on the device they are on, labels of the range are routed
    there; a label outside every range is routed by the modulo rule."""
    import torch
    n, dim, k = 50000, 32, 10
    X = make_vectors(0, n, dim, seed=3, dist="normal")
    Q = make_vectors(0, 16, dim, seed=4, dist="normal")
    have = capi.device_count()
    devs = [0, 1 % have]
    capi.configure_devices(devs)
    try:
        G = capi.BFIndex(capi.BFParams(type=0, dim=dim, metric=1, multi=False, initialCapacity=n, blockSize=1024))
        half = n // 2
        for part, (a, b) in enumerate(((0, half), (half, n))):
            t = torch.from_numpy(X[a:b]).to("cuda:%d" % devs[part])
            torch.cuda.synchronize()
            assert G.add_device_rows(t.data_ptr(), dim * 4, b - a, a) == b - a
        assert G.index_size() == n
        G.add_vector(X[5] * 2, n + 3)                                       # modulo-routed label
        P = port.PortIndex(0, dim, 1)
        P.add_many(X)
        P.add(X[5] * 2, n + 3)
        labels, scores = G.knn_batch(Q, k)
        for i in range(len(Q)):
            pl, ps, _ = P.topk(Q[i], k)
            assert np.array_equal(labels[i], pl.astype(np.int64)) and np.array_equal(scores[i], ps), i
        assert G.delete_vector(half + 1) == P.delete(half + 1) == 1        # a label inside the second range
        l, s = G.knn_query(X[half + 1], 3)
        pl, ps, _ = P.topk(X[half + 1], 3)
        assert np.array_equal(l[0], pl.astype(np.int64))
        G.close()
        P.close()
    finally:
        capi.set_device(0)
