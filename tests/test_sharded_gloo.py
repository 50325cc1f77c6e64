"""The N>1 plumbing on CPU: world_size-2 gloo processes all-gather their per-shard top-k lists and merge
them; the merged lists must equal the oracle's answer over the whole (unsharded) index. The per-shard
top-k here comes from the oracle port (no GPU in this container) — what is under test is the sharding
arithmetic, the collective layout and the merge order of vectorsimilarity_b200/sharded.py."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port_no, n, dim, k, nq, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from datagen import make_vectors
    from oracle import port
    from vectorsimilarity_b200 import sharded
    X = make_vectors(4, n, dim, seed=1)            # int8 L2: plenty of exact ties across shards
    X = (X // 32).astype(np.int8)
    Q = (make_vectors(4, nq, dim, seed=2) // 32).astype(np.int8)
    lo, hi = sharded.shard_bounds(n, world, rank)
    P = port.PortIndex(4, dim, 0)
    P.add_many(X[lo:hi], first_label=lo)
    ls = np.full((nq, k), np.nan, dtype=np.float32)
    ll = np.full((nq, k), np.iinfo(np.uint64).max, dtype=np.uint64)
    for i in range(nq):
        l, s, _ = P.topk(Q[i], k)
        ll[i, :len(l)], ls[i, :len(l)] = l, s
    ms, ml = sharded.gather_merge_host(ls, ll, k)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), s=ms, l=ml)
    dist.destroy_process_group()


def test_two_rank_gather_merge(tmp_path, port):
    n, dim, k, nq, world = 501, 6, 20, 7, 2
    port_no = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port_no, n, dim, k, nq, str(tmp_path)), nprocs=world, join=True)
    from datagen import make_vectors
    X = (make_vectors(4, n, dim, seed=1) // 32).astype(np.int8)
    Q = (make_vectors(4, nq, dim, seed=2) // 32).astype(np.int8)
    P = port.PortIndex(4, dim, 0)
    P.add_many(X)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["l"], r1["l"]) and np.array_equal(r0["s"], r1["s"])      # every rank holds the answer
    for i in range(nq):
        l, s, _ = P.topk(Q[i], k)
        assert np.array_equal(r0["l"][i], l)
        assert np.array_equal(r0["s"][i].astype(np.float64), s)


def test_shard_bounds_cover_everything():
    from vectorsimilarity_b200 import sharded
    for n in (0, 1, 7, 1000, 10_000_000):
        for w in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
