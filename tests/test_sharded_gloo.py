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


class _FakeLocalIndex:
    """Stands in for capi.BFIndex on CPU (same return shapes): exact L2 over this rank's rows."""

    def __init__(self, X, first_label):
        self.X, self.first = X.astype(np.float64), first_label

    def _all(self, q):
        s = ((self.X - q.astype(np.float64)) ** 2).sum(1)
        l = np.arange(self.first, self.first + len(s), dtype=np.int64)
        o = np.lexsort((l, s))
        return l[o], s[o]

    def range_query(self, q, radius, order=0):
        l, s = self._all(q)
        keep = s <= radius
        return l[keep].reshape(1, -1), s[keep].reshape(1, -1)

    def create_batch_iterator(self, q):
        outer = self

        class It:
            def __init__(self):
                self.l, self.s = outer._all(q)
                self.pos = 0

            def has_next(self):
                return self.pos < len(self.l)

            def get_next_results(self, n, order=0):
                a, self.pos = self.pos, min(self.pos + n, len(self.l))
                return self.l[a:self.pos].reshape(1, -1), self.s[a:self.pos].reshape(1, -1)

            def reset(self):
                self.pos = 0

            def close(self):
                pass
        return It()


def _worker_varlen(rank, world, port_no, n, dim, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vectorsimilarity_b200 import sharded
    rng = np.random.default_rng(3)
    X = rng.integers(-3, 4, (n, dim)).astype(np.float32)      # small integers: exact ties across shards
    q = rng.integers(-3, 4, dim).astype(np.float32)
    lo, hi = sharded.shard_bounds(n, world, rank)
    S = sharded.ShardedFlatIndex.__new__(sharded.ShardedFlatIndex)   # collective logic only: no device index here
    S.local, S.group, S.device = _FakeLocalIndex(X[lo:hi], lo), None, None
    out = {}
    for name, radius in (("r_small", 6.0), ("r_none", -1.0), ("r_all", 1e9)):
        for oname, order in (("score", 0), ("id", 1)):
            l, s = S.range_query(q, radius, order=order)
            out[f"{name}_{oname}_l"], out[f"{name}_{oname}_s"] = l, s
    it = S.create_batch_iterator(q)
    batches, sizes = [], [1, 7, 50, 3, 200, 1000]
    i = 0
    while it.has_next() and i < 50:
        l, s = it.get_next_results(sizes[i % len(sizes)])
        batches.append((l[0], s[0]))
        i += 1
    out["it_l"] = np.concatenate([b[0] for b in batches])
    out["it_s"] = np.concatenate([b[1] for b in batches])
    out["it_sizes"] = np.array([len(b[0]) for b in batches])
    it.reset()
    l, s = it.get_next_results(5, order=1)
    out["it_reset_l"] = l[0]
    np.savez(os.path.join(out_dir, "v%d.npz" % rank), **out)
    dist.destroy_process_group()


def test_two_rank_range_and_batch_iterator(tmp_path):
    """Sharded range query and batch iterator: variable-length all-gather + merge equals one index over
    all rows — by score, by id, empty and everything; iterator batches of uneven sizes cover every label once in
    ascending (score, label) order."""
    n, dim, world = 403, 5, 2
    port_no = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_varlen, args=(world, port_no, n, dim, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    X = rng.integers(-3, 4, (n, dim)).astype(np.float32)
    q = rng.integers(-3, 4, dim).astype(np.float32)
    whole = _FakeLocalIndex(X, 0)
    al, as_ = whole._all(q)
    r0, r1 = np.load(tmp_path / "v0.npz"), np.load(tmp_path / "v1.npz")
    for key in r0.files:
        assert np.array_equal(r0[key], r1[key]), key
    for name, radius in (("r_small", 6.0), ("r_none", -1.0), ("r_all", 1e9)):
        keep = as_ <= radius
        assert np.array_equal(r0[name + "_score_l"], al[keep]) and np.array_equal(r0[name + "_score_s"], as_[keep])
        o = np.argsort(al[keep], kind="stable")
        assert np.array_equal(r0[name + "_id_l"], al[keep][o]) and np.array_equal(r0[name + "_id_s"], as_[keep][o])
    assert np.array_equal(r0["it_l"], al) and np.array_equal(r0["it_s"], as_)
    assert r0["it_sizes"].tolist()[:5] == [1, 7, 50, 3, 200] and r0["it_sizes"].sum() == n
    assert np.array_equal(r0["it_reset_l"], np.sort(al[:5]))
