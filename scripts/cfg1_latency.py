"""BASELINE configs[0]: flat fp32 L2, N=100k, d=128, K=10, single query. Latency of VecSimIndex_TopKQuery through the C API
(host blob in, reply object out) on one B200, next to the unmodified reference's topKQuery on one host core, on the same
rows and queries; every id and score is compared.

    python scripts/cfg1_latency.py > gpurun_out/cfg1_latency.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vectorsimilarity_b200 import build, capi  # noqa: E402

N, DIM, K, NQ = 100_000, 128, 10, 300


def main():
    build.build()
    rng = np.random.default_rng(47)
    X = rng.uniform(-1, 1, (N, DIM)).astype(np.float32)
    Q = rng.uniform(-1, 1, (NQ, DIM)).astype(np.float32)
    G = capi.BFIndex(capi.BFParams(type=0, dim=DIM, metric=0, multi=False, initialCapacity=N, blockSize=1024))
    G.add_vectors(X)
    for q in Q[:20]:
        G.knn_query(q, K)
    lat, res = [], []
    for q in Q:
        t0 = time.perf_counter()
        l, s = G.knn_query(q, K)
        lat.append(time.perf_counter() - t0)
        res.append((l[0].copy(), s[0].copy()))
    st = G.last_query_stats()
    out = {"workload": "flat_fp32_l2_100k_d128_k10_b1", "queries": NQ,
           "gpu_latency_us": {"median": float(np.median(lat) * 1e6), "p10": float(np.percentile(lat, 10) * 1e6),
                              "p99": float(np.percentile(lat, 99) * 1e6)},
           "gpu_device_ms_last_query": st["total_ms"], "gpu_kernel_launches_per_query": st["kernel_launches"]}
    from oracle import ref
    if ref.available():
        R = ref.RefIndex(0, DIM, 0)
        R.add_many(X)
        rl = []
        same = True
        for i, q in enumerate(Q):
            t0 = time.perf_counter()
            l, s, _ = R.topk(q, K)
            rl.append(time.perf_counter() - t0)
            same &= bool(np.array_equal(l.astype(np.int64), res[i][0]) and np.array_equal(s, res[i][1]))
        out["reference_1core_latency_us"] = {"median": float(np.median(rl) * 1e6), "p99": float(np.percentile(rl, 99) * 1e6)}
        out["identical_to_reference"] = same
    print(json.dumps(out))


if __name__ == "__main__":
    main()
