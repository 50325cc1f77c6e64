"""HNSW (BASELINE.json configs[4] shape: fp32 L2 d=128 M=16 efC=200 efR=64 K=10 batch=256) on one B200:
device build time, batched search QPS, distance evaluations and the bytes they touch, next to the
unmodified reference (oracle/_ref) on the host cores over the same vectors. Not the headline bench
(bench.py measures configs[1]); prints one JSON line.

    python scripts/hnsw_bench.py --rows 100000 [--ref]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=100000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--M", type=int, default=16)
    ap.add_argument("--efc", type=int, default=200)
    ap.add_argument("--ef", type=int, default=64)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--ref", action="store_true", help="also build + search with the unmodified reference (CPU)")
    a = ap.parse_args()
    from vectorsimilarity_b200 import build, capi
    build.build()
    rng = np.random.default_rng(47)
    X = rng.uniform(-1, 1, (a.rows, a.dim)).astype(np.float32)
    Q = rng.uniform(-1, 1, (a.batch, a.dim)).astype(np.float32)
    G = capi.HNSWIndex(capi.HNSWParams(type=0, dim=a.dim, metric=0, multi=False, initialCapacity=a.rows, blockSize=1024,
                                       M=a.M, efConstruction=a.efc, efRuntime=a.ef, epsilon=0.01))
    t0 = time.perf_counter()
    G.add_vectors(X)
    G.knn_batch(Q[:1], a.k)  # flushes the pending inserts
    build_s = time.perf_counter() - t0
    out = {"workload": f"hnsw_fp32_l2_{a.rows}_d{a.dim}_M{a.M}_efc{a.efc}_ef{a.ef}_k{a.k}_b{a.batch}",
           "build_s": build_s, "build_us_per_insert": build_s / a.rows * 1e6}
    for _ in range(3):
        labels, scores = G.knn_batch(Q, a.k)
    ms, wall = [], []
    for _ in range(a.steps):
        t0 = time.perf_counter()
        labels, scores = G.knn_batch(Q, a.k)
        wall.append(time.perf_counter() - t0)
        st = G.hnsw_stats()
        ms.append(st["ms"])
    row_bytes = a.dim * 4
    evals = st["dist_evals"]
    hops = st["hops"]
    bytes_touched = evals * (row_bytes + 4 + 4) + hops * (2 * a.M + 1) * 4
    kms = float(np.median(ms))
    out.update({"search_kernel_ms": kms, "search_qps_device": a.batch / (kms * 1e-3),
                "search_qps_e2e": a.batch / float(np.median(wall)),
                "dist_evals_per_query": evals / a.batch, "hops_per_query": hops / a.batch,
                "achieved_gbs": bytes_touched / (kms * 1e-3) / 1e9})
    # recall against the exact flat scan on the device
    F = capi.BFIndex(capi.BFParams(type=0, dim=a.dim, metric=0, multi=False, initialCapacity=a.rows, blockSize=1024))
    F.add_vectors(X)
    fl, _ = F.knn_batch(Q, a.k)
    out["recall_at_k"] = float(np.mean([len(set(labels[i]) & set(fl[i])) / a.k for i in range(a.batch)]))
    F.close()
    # throughput at a larger batch (several waves of CTAs: the tail of the slowest query no longer dominates)
    if a.batch < 4096:
        Qb = rng.uniform(-1, 1, (4096, a.dim)).astype(np.float32)
        G.knn_batch(Qb, a.k)
        G.knn_batch(Qb, a.k)
        out["search_qps_device_b4096"] = 4096 / (G.hnsw_stats()["ms"] * 1e-3)
    if a.ref:
        from oracle import ref
        ref.lib()
        R = ref.RefIndex(0, a.dim, 0, algo="hnsw", M=a.M, ef_construction=a.efc, ef_runtime=a.ef)
        t0 = time.perf_counter()
        R.add_many(X)
        out["ref_build_s"] = time.perf_counter() - t0
        threads = os.cpu_count() or 1
        rl, rs, secs = R.topk_many(Q, a.k, n_threads=threads, ef_runtime=a.ef)
        rl, rs, secs = R.topk_many(Q, a.k, n_threads=threads, ef_runtime=a.ef)
        out["ref_qps_all_cores"] = a.batch / secs
        out["ref_cores"] = threads
        _, _, secs1 = R.topk_many(Q, a.k, n_threads=1, ef_runtime=a.ef)
        out["ref_qps_1core"] = a.batch / secs1
        out["ids_identical_to_reference"] = bool(np.array_equal(rl.astype(np.int64), labels))
        out["scores_identical_to_reference"] = bool(np.array_equal(rs, scores))
        R.close()
    G.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
