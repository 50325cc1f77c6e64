"""Generates tests/golden/hnsw_case.npz from the UNMODIFIED reference (oracle/_ref): vectors, the
graph the reference's single-threaded HNSWIndex builds over them (levels, links per level, entry
point), and what its topKQuery / rangeQuery return. The GPU tests check (a) the device builder
produces this exact graph from the vectors alone, (b) the device traversal returns these exact
results on it.

    python tests/golden/make_hnsw_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
from datagen import make_vectors  # noqa: E402

N, DIM, M, EFC, NQ, K = 1500, 24, 8, 48, 16, 10


def main():
    ref.lib()
    out = {}
    for name, metric in (("l2", 0), ("cos", 2)):
        X = make_vectors(0, N, DIM, seed=77 + metric)
        Q = make_vectors(0, NQ, DIM, seed=78 + metric)
        idx = ref.RefIndex(0, DIM, metric, algo="hnsw", M=M, ef_construction=EFC, ef_runtime=10)
        idx.add_many(X)
        g = idx.hnsw_export()
        out[name + "_X"], out[name + "_Q"] = X, Q
        out[name + "_stored"] = g["vectors"]
        out[name + "_levels"] = g["levels"]
        out[name + "_entry"] = np.array([g["entry"], g["max_level"]])
        for lvl in range(len(g["links"])):
            out[f"{name}_links{lvl}"] = g["links"][lvl]
            out[f"{name}_counts{lvl}"] = g["counts"][lvl]
        for ef in (10, 40):
            labels = np.zeros((NQ, K), dtype=np.int64)
            scores = np.zeros((NQ, K), dtype=np.float64)
            for i in range(NQ):
                l, s, code = idx.topk(Q[i], K, ef_runtime=ef)
                assert code == 0 and len(l) == K
                labels[i], scores[i] = l, s
            out[f"{name}_labels_ef{ef}"], out[f"{name}_scores_ef{ef}"] = labels, scores
        radii, rl_all, rs_all, rn = [], [], [], []
        for i in range(NQ):
            radius = float(out[f"{name}_scores_ef40"][i][5])
            l, s, _ = idx.range(Q[i], radius)
            order = np.lexsort((l, s))
            radii.append(radius)
            rn.append(len(l))
            rl_all.append(l[order].astype(np.int64))
            rs_all.append(s[order])
        out[name + "_radius"] = np.array(radii)
        out[name + "_range_n"] = np.array(rn)
        out[name + "_range_labels"] = np.concatenate(rl_all)
        out[name + "_range_scores"] = np.concatenate(rs_all)
        idx.close()
    np.savez_compressed(os.path.join(HERE, "hnsw_case.npz"), **out)
    print("wrote hnsw_case.npz;", len(out), "arrays,", os.path.getsize(os.path.join(HERE, "hnsw_case.npz")), "bytes")


if __name__ == "__main__":
    main()
