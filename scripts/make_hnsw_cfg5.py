"""One-off (build container, CPU): build BASELINE.json configs[4]'s graph (fp32 L2, N=1M, d=128, M=16, efC=200) with the
UNMODIFIED reference, export it in the vsgpu_hnsw_import layout and record the reference's answers to 256 queries
(efR=64, K=10). The file (hnsw_cache/, git-ignored) travels to the GPU box with the snapshot, where
scripts/hnsw_cfg5_check.py loads the graph onto the device, checks ids/scores against the recorded answers and
times the search. Vectors are regenerated from the seed on both sides."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402

N = int(os.environ.get("CFG5_ROWS", 1_000_000))
DIM, M, EFC, EF, K, NQ = 128, 16, 200, 64, 10, 256


def vectors():
    rng = np.random.default_rng(47)
    return rng.uniform(-1, 1, (N, DIM)).astype(np.float32), rng.uniform(-1, 1, (NQ, DIM)).astype(np.float32)


def main():
    X, Q = vectors()
    ref.lib()
    R = ref.RefIndex(0, DIM, 0, algo="hnsw", M=M, ef_construction=EFC, ef_runtime=EF)
    t0 = time.perf_counter()
    step = 50_000
    for a in range(0, N, step):
        R.add_many(X[a:a + step], first_label=a)
        print("built %d / %d in %.0f s" % (min(a + step, N), N, time.perf_counter() - t0), flush=True)
    build_s = time.perf_counter() - t0
    g = R.hnsw_export()
    levels = g["levels"]
    l0 = np.zeros((N, 2 * M + 1), dtype=np.uint32)
    l0[:, 0] = g["counts"][0]
    l0[:, 1:] = np.where(np.arange(2 * M)[None, :] < g["counts"][0][:, None], g["links"][0], 0)
    recs = []
    for i in np.nonzero(levels)[0]:
        for lvl in range(1, int(levels[i]) + 1):
            r = np.zeros(M + 1, dtype=np.uint32)
            c = int(g["counts"][lvl][i])
            r[0] = c
            r[1:1 + c] = g["links"][lvl][i][:c]
            recs.append(r)
    upper = np.stack(recs) if recs else np.zeros((0, M + 1), dtype=np.uint32)
    labels = np.zeros((NQ, K), dtype=np.int64)
    scores = np.zeros((NQ, K), dtype=np.float64)
    t0 = time.perf_counter()
    for i in range(NQ):
        l, s, _ = R.topk(Q[i], K, ef_runtime=EF)
        labels[i], scores[i] = l, s
    q_s = time.perf_counter() - t0
    np.savez(os.path.join(ROOT, "hnsw_cache", "cfg5_graph_%d.npz" % N), levels=levels, l0=l0, upper=upper,
             entry=np.array([g["entry"], g["max_level"]]), ref_labels=labels, ref_scores=scores,
             ref_build_s=np.array(build_s), ref_query_s_1core=np.array(q_s))
    print("done: build %.0f s (%.0f us/insert), %d upper records, 1-core search %.1f q/s" %
          (build_s, build_s / N * 1e6, len(upper), NQ / q_s))


if __name__ == "__main__":
    main()
