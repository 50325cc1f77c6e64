"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs; not collected by pytest):
the sharded flat index (one rank per GPU, NCCL all-gather of per-shard top-K + device merge) must return
exactly what the oracle returns over the whole, unsharded row set — ids, order and scores.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/nccl_parity_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from datagen import make_vectors
    from oracle import port
    from vectorsimilarity_b200 import build, capi, sharded
    if rank == 0:
        build.build()
    dist.barrier()
    capi.lib()
    capi.set_device(local)
    port.set_tier(port.TIER_AVX512)
    checked = 0
    for vtype, metric, n, dim, k, nq, mode in [(0, 1, 20011, 96, 25, 40, 0), (0, 0, 5003, 128, 10, 3, 1), (4, 0, 3001, 6, 20, 9, 1),
                                               (2, 1, 30011, 64, 50, 64, 0), (1, 2, 1001, 24, 7, 5, 1),
                                               # every shard large enough for the tensor path (forced): the phased call with
                                               # its bound exchange after every phase
                                               (0, 1, 40000 * world + 11, 64, 100, 96, 2), (0, 0, 36000 * world, 72, 10, 33, 2),
                                               (2, 2, 34000 * world + 5, 64, 37, 24, 2), (4, 2, 40000 * world, 64, 10, 64, 2)]:
        X = make_vectors(vtype, n, dim, seed=11 + vtype)
        Q = make_vectors(vtype, nq, dim, seed=12 + vtype)
        if vtype == 4:  # coarse int8: many exact ties across the shard boundary
            X, Q = (X // 32).astype(np.int8), (Q // 32).astype(np.int8)
        lo, hi = sharded.shard_bounds(n, world, rank)
        params = capi.BFParams(type=vtype, dim=dim, metric=metric, multi=False, initialCapacity=hi - lo, blockSize=1024)
        S = sharded.ShardedFlatIndex(params)
        S.add_vectors(X[lo:hi], labels=np.arange(lo, hi, dtype=np.uint64))
        P = port.PortIndex(vtype, dim, metric)
        P.add_many(X)
        # processed queries (cosine: normalised like the index does)
        Qp = Q.copy()
        if metric == 2 and vtype >= 4:   # int8 / uint8 cosine blobs carry their norm after the dim bytes
            Qp = np.zeros((nq, dim + 4), dtype=Q.dtype)
            Qp[:, :dim] = Q
            for q in Qp:
                port.normalize(vtype, dim, q)
        elif metric == 2:
            Qp = np.stack([port.normalize(vtype, dim, q.copy()) for q in Qp])
        labels, scores = S.knn_batch(Qp, k, flags=mode)
        for i in range(nq):
            pl, ps, _ = P.topk(Q[i], k)
            assert np.array_equal(labels[i].view(np.uint64), pl), (rank, vtype, metric, i, labels[i], pl)
            assert np.array_equal(scores[i], ps), (rank, vtype, metric, i)
            checked += 1
        # range query and batch iterator across the shards (variable-length all-gather)
        pl, ps, _ = P.topk(Q[0], min(60, n))
        radius = float(ps[-1]) if ps[-1] >= 0 else 0.25   # IP scores of unnormalised rows go negative; radius may not
        for order in (capi.BY_SCORE, capi.BY_ID):
            wl, ws, _ = P.range(Q[0], radius, order=order)
            gl, gs = S.range_query(Q[0], radius, order=order)
            assert np.array_equal(gl.view(np.uint64), wl) and np.array_equal(gs, ws), (rank, vtype, metric, "range", order)
            checked += 1
        it, pit = S.create_batch_iterator(Q[0]), P.batch_iterator(Q[0])
        for bs in (1, 10, 37, 100):
            gl, gs = it.get_next_results(bs)
            wl, ws, _ = pit.next(bs)
            assert np.array_equal(gl[0].view(np.uint64), wl) and np.array_equal(gs[0], ws), (rank, vtype, metric, "iterator", bs)
            checked += 1
        it.reset()
        pit.reset()
        gl, gs = it.get_next_results(5)
        wl, ws, _ = pit.next(5)
        assert np.array_equal(gl[0].view(np.uint64), wl) and np.array_equal(gs[0], ws) and it.has_next()
        it.close()
        pit.close()
        S.close()
        P.close()
    t = torch.tensor([checked], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("nccl parity OK: %d query results on %d ranks identical to the unsharded oracle" % (int(t.item()), world))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
