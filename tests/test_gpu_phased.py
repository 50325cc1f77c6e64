"""Phased sharded top-k (vsgpu_topk_device_begin / _next / _finish): several shards of one row set, here all on
one GPU with one store each, run their coarse phases in lock step and exchange their bounds after every phase the way the NCCL
front-end does (element-wise max over the shards' buffers). Merged, the shards' lists must be exactly what the oracle returns
over the unsharded rows — ids, order and fp32 scores bit for bit — although every shard now admits and re-ranks against a
bound that comes from the OTHER shards' rows."""
import ctypes as C

import numpy as np
import pytest

from datagen import make_vectors

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch
    from oracle import port
    from vectorsimilarity_b200 import build, capi, sharded
    build.build()
    capi.lib()
    port.set_tier(port.TIER_AVX512)
    yield {"torch": torch, "capi": capi, "port": port, "sharded": sharded, "G": sharded._vsgpu()}
    capi.set_topk_mode(0)


def _run_phased(env, stores, Qp, k, flags, rounds, skip_next=0, f64=False):
    torch, G = env["torch"], env["G"]
    W, nq = len(stores), Qp.shape[0]
    dev = torch.device("cuda", 0)
    q_dev = torch.from_numpy(np.ascontiguousarray(Qp).view(np.uint8).reshape(nq, -1)).to(dev)
    torch.cuda.synchronize()
    bounds = [torch.empty(2 * nq, dtype=torch.float32, device=dev) for _ in range(W)]
    scores = [torch.empty((nq, k), dtype=torch.float64 if f64 else torch.float32, device=dev) for _ in range(W)]
    labels = [torch.empty((nq, k), dtype=torch.int64, device=dev) for _ in range(W)]

    def exchange():
        for st in stores:
            assert G.vsgpu_store_sync(st) == 0
        red = torch.stack(bounds).max(dim=0).values
        for b in bounds:
            b.copy_(red)
        torch.cuda.synchronize()
        return red

    for i, st in enumerate(stores):
        rc = G.vsgpu_topk_device_begin(st, q_dev.data_ptr(), nq, q_dev.stride(0), k, flags, labels[i].data_ptr(), scores[i].data_ptr(),
                                       None, W, rounds, bounds[i].data_ptr())
        assert rc == 0, G.vsgpu_last_error()
    history = []
    for _ in range(rounds - 1 - skip_next):
        history.append(exchange())
        for i, st in enumerate(stores):
            assert G.vsgpu_topk_device_next(st, bounds[i].data_ptr()) == 0, G.vsgpu_last_error()
    history.append(exchange())
    for i, st in enumerate(stores):
        assert G.vsgpu_topk_device_finish(st, bounds[i].data_ptr()) == 0, G.vsgpu_last_error()
    stats = []
    for st in stores:
        assert G.vsgpu_store_sync(st) == 0, G.vsgpu_last_error()
        s = env["sharded"].Stats()
        G.vsgpu_last_stats(st, C.byref(s))
        stats.append(s)
    S = torch.stack(scores).cpu().numpy()
    L = torch.stack(labels).cpu().numpy()
    return S, L, stats, [h.cpu().numpy() for h in history]


@pytest.mark.parametrize("vtype,metric,n,dim,k,nq,W", [
    (0, 1, 400_000, 64, 100, 64, 8),     # the headline's shape in small: 8 shards, k = 100
    (0, 0, 300_011, 72, 10, 40, 3),      # L2, shards of unequal size
    (2, 2, 270_000, 64, 37, 24, 4),      # bf16 cosine
    (1, 1, 140_000, 32, 5, 9, 2),        # fp64 stores never take the tensor path: _begin does all the work
    (0, 1, 262_144, 64, 500, 16, 2),     # the large-k buffers
    (3, 1, 200_000, 80, 100, 33, 5),     # fp16 rows, k / shards = 20
])
def test_phased_shards_equal_unsharded_oracle(env, vtype, metric, n, dim, k, nq, W):
    capi, port, sharded, G = env["capi"], env["port"], env["sharded"], env["G"]
    X = make_vectors(vtype, n, dim, seed=31 + vtype, dist="normal")
    Q = make_vectors(vtype, nq, dim, seed=32 + vtype, dist="normal")
    idx, stores = [], []
    for r in range(W):
        lo, hi = sharded.shard_bounds(n, W, r)
        I = capi.BFIndex(capi.BFParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=hi - lo, blockSize=1024))
        I.add_vectors(X[lo:hi], labels=np.arange(lo, hi, dtype=np.uint64))
        idx.append(I)
        stores.append(I.device_store())
    P = port.PortIndex(vtype, dim, metric)
    P.add_many(X)
    Qp = Q.copy()
    if metric == 2:
        Qp = np.stack([port.normalize(vtype, dim, q.copy()) for q in Qp])
    rounds = int(G.vsgpu_topk_rounds((n + W - 1) // W, k, W))
    assert 1 <= rounds <= 6
    flags = 0 if vtype == 1 else 2                      # tensor path forced where the type has one
    hist0 = None
    for skip in (0, 1 if rounds > 1 else 0):            # a caller that exchanges less often than agreed still gets the result
        S, L, stats, hist = _run_phased(env, stores, Qp, k, flags, rounds, skip_next=skip, f64=vtype == 1)
        hist0 = hist0 or hist
        if vtype != 1:
            assert all(s.path == 1 for s in stats) and sum(s.fallback_queries for s in stats) == 0
        ms, ml = sharded.merge_topk_host(S.astype(np.float64), L, k)
        for i in range(nq):
            pl, ps, _ = P.topk(Q[i], k)
            assert np.array_equal(ml[i].view(np.uint64), pl), (i, skip)
            assert np.array_equal(ms[i], ps if vtype == 1 else ps.astype(np.float32).astype(np.float64)), (i, skip)
        if vtype != 1 and skip == 0:
            # the exchanged bound is a lower bound of the true k-th score in the admission test's units and never loosens
            comb = [np.maximum(h[:nq], -h[nq:]) for h in hist0]
            for a, b in zip(comb, comb[1:]):
                assert np.all(b >= a)
            assert np.all(np.isfinite(comb[-1]))
    # the second value is what makes the bound global: with 8 shards the shards' own k-th best is far below the k-th overall
    if W == 8:
        own_kth = hist0[-1][:nq]
        assert np.mean(comb[-1] > own_kth) > 0.9
    for I in idx:
        I.close()
    P.close()


def test_phased_candidates_shrink_with_the_exchange(env):
    """What the exchange buys: with 8 shards the rows admitted per query (and so the merge and re-rank work) must be well
    below the unshared run's, for identical results."""
    capi, sharded, G = env["capi"], env["sharded"], env["G"]
    n, dim, k, nq, W = 8 * 150_000, 64, 100, 128, 8
    X = make_vectors(0, n, dim, seed=41, dist="normal")
    Q = make_vectors(0, nq, dim, seed=42, dist="normal")
    idx = []
    for r in range(W):
        lo, hi = sharded.shard_bounds(n, W, r)
        I = capi.BFIndex(capi.BFParams(type=0, dim=dim, metric=1, multi=False, initialCapacity=hi - lo, blockSize=1024))
        I.add_vectors(X[lo:hi], labels=np.arange(lo, hi, dtype=np.uint64))
        idx.append(I)
    stores = [I.device_store() for I in idx]
    rounds = int(G.vsgpu_topk_rounds(n // W, k, W))
    S1, L1, st1, _ = _run_phased(env, stores, Q, k, 2, rounds)
    # the same shards, each on its own (plain call)
    torch = env["torch"]
    dev = torch.device("cuda", 0)
    q_dev = torch.from_numpy(Q.view(np.uint8).reshape(nq, -1)).to(dev)
    sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    lb = torch.empty((nq, k), dtype=torch.int64, device=dev)
    own_cands, S0, L0 = 0, [], []
    for st in stores:
        assert G.vsgpu_topk_device(st, q_dev.data_ptr(), nq, q_dev.stride(0), k, 2, lb.data_ptr(), sc.data_ptr(), None) == 0
        assert G.vsgpu_store_sync(st) == 0
        s = sharded.Stats()
        G.vsgpu_last_stats(st, C.byref(s))
        own_cands += s.candidates
        S0.append(sc.cpu().numpy().copy())
        L0.append(lb.cpu().numpy().copy())
    shared_cands = sum(s.candidates for s in st1)
    assert shared_cands < 0.85 * own_cands, (shared_cands, own_cands)
    a = sharded.merge_topk_host(S1.astype(np.float64), L1, k)
    b = sharded.merge_topk_host(np.stack(S0).astype(np.float64), np.stack(L0), k)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    for I in idx:
        I.close()
