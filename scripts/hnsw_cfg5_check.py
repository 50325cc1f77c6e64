"""BASELINE.json configs[4] on one B200: load the graph the UNMODIFIED reference built (hnsw_cache/cfg5_graph_<N>.npz,
made by scripts/make_hnsw_cfg5.py on the CPU), search it on the device with efR=64, K=10, batch=256, check every id
and score against the reference's recorded answers, and time the traversal. Prints one JSON line."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N = int(os.environ.get("CFG5_ROWS", 1_000_000))
DIM, M, EFC, EF, K, NQ = 128, 16, 200, 64, 10, 256


def main():
    from vectorsimilarity_b200 import build, capi
    build.build()
    g = np.load(os.path.join(ROOT, "hnsw_cache", "cfg5_graph_%d.npz" % N))
    rng = np.random.default_rng(47)
    X = rng.uniform(-1, 1, (N, DIM)).astype(np.float32)
    Q = rng.uniform(-1, 1, (NQ, DIM)).astype(np.float32)
    G = capi.HNSWIndex(capi.HNSWParams(type=0, dim=DIM, metric=0, multi=False, initialCapacity=N, blockSize=1024, M=M,
                                       efConstruction=EFC, efRuntime=EF, epsilon=0.01))
    levels, l0, upper = (np.ascontiguousarray(g[k]) for k in ("levels", "l0", "upper"))
    t0 = time.perf_counter()
    rc = capi.lib().VecSimGPU_HNSWImportGraph(G._h, X.ctypes.data, 1, N, None, levels.ctypes.data, l0.ctypes.data,
                                              upper.ctypes.data if len(upper) else None, len(upper), int(g["entry"][0]),
                                              int(g["entry"][1]))
    assert rc == 0, capi.lib().VecSimGPU_LastError()
    load_s = time.perf_counter() - t0
    for _ in range(3):
        labels, scores = G.knn_batch(Q, K)
    ms, wall = [], []
    for _ in range(5):
        t0 = time.perf_counter()
        labels, scores = G.knn_batch(Q, K)
        wall.append(time.perf_counter() - t0)
        st = G.hnsw_stats()
        ms.append(st["ms"])
    kms = float(np.median(ms))
    evals, hops = st["dist_evals"], st["hops"]
    bytes_touched = evals * (DIM * 4 + 4 + 1 + 4) + hops * (2 * M + 1) * 4
    out = {"workload": "hnsw_fp32_l2_%d_d128_M16_efc200_ef64_k10_b256" % N, "graph": "built by the unmodified reference",
           "ids_identical_to_reference": bool(np.array_equal(labels, g["ref_labels"])),
           "scores_identical_to_reference": bool(np.array_equal(scores, g["ref_scores"])),
           "search_kernel_ms": kms, "search_qps_device": NQ / (kms * 1e-3), "search_qps_e2e": NQ / float(np.median(wall)),
           "dist_evals_per_query": evals / NQ, "hops_per_query": hops / NQ, "achieved_gbs": bytes_touched / (kms * 1e-3) / 1e9,
           "bytes_per_eval": DIM * 4 + 9, "graph_load_s": load_s,
           "ref_build_s_cpu": float(g["ref_build_s"]), "ref_qps_1core_build_host": NQ / float(g["ref_query_s_1core"])}
    Qb = rng.uniform(-1, 1, (4096, DIM)).astype(np.float32)
    G.knn_batch(Qb, K)
    G.knn_batch(Qb, K)
    out["search_qps_device_b4096"] = 4096 / (G.hnsw_stats()["ms"] * 1e-3)
    try:  # the reference on this box's cores, over the same vectors, needs its own build: too slow here; flat recall instead
        F = capi.BFIndex(capi.BFParams(type=0, dim=DIM, metric=0, multi=False, initialCapacity=N, blockSize=1024))
        F.add_vectors(X)
        fl, _ = F.knn_batch(Q, K)
        out["recall_at_k_vs_flat"] = float(np.mean([len(set(labels[i]) & set(fl[i])) / K for i in range(NQ)]))
        F.close()
    except Exception as e:
        out["recall_at_k_vs_flat"] = repr(e)
    G.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
