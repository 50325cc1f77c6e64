"""Generates tests/golden/flat_cases.npz from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
The fixture stores inputs AND the reference's outputs, so the GPU box (which has neither
/root/reference nor necessarily the same CPU) can check both the oracle port and the CUDA path
against what the real reference returned. avx512_fp16 is masked.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref  # noqa: E402
from datagen import make_vectors, TYPE_NAMES, METRIC_NAMES  # noqa: E402

N, DIM, K, NQ = 400, 48, 10, 3


def main():
    ref.lib()
    ref.set_disabled_features("avx512_fp16")
    out = {"host_features": np.array(" ".join(ref.host_features()))}
    for vtype in range(6):
        for metric in range(3):
            name = f"{TYPE_NAMES[vtype]}_{METRIC_NAMES[metric]}"
            X = make_vectors(vtype, N, DIM, seed=1000 + vtype * 10 + metric)
            Q = make_vectors(vtype, NQ, DIM, seed=2000 + vtype * 10 + metric)
            idx = ref.RefIndex(vtype, DIM, metric, block_size=64)
            idx.add_many(X)
            labels = np.zeros((NQ, K), dtype=np.int64)
            scores = np.zeros((NQ, K), dtype=np.float64)
            rng_counts = np.zeros(NQ, dtype=np.int64)
            radii = np.zeros(NQ, dtype=np.float64)
            for i in range(NQ):
                l, s, code = idx.topk(Q[i], K)
                assert code == 0 and len(l) == K
                labels[i], scores[i] = l, s
                radii[i] = max(float(s[-1]), 0.0)
                rl, rs, _ = idx.range(Q[i], radii[i])
                rng_counts[i] = len(rl)
            idx.close()
            out[name + "_X"], out[name + "_Q"] = X, Q
            out[name + "_labels"], out[name + "_scores"] = labels, scores
            out[name + "_radius"], out[name + "_range_count"] = radii, rng_counts
    # tie case: int8 L2, tiny value range, shuffled labels, k crossing a tie group
    rng = np.random.default_rng(5)
    X = rng.integers(-2, 3, (300, 3)).astype(np.int8)
    lab = rng.permutation(1000)[:300].astype(np.uint64)
    q = np.array([0, 1, -1], dtype=np.int8)
    idx = ref.RefIndex(4, 3, 0, block_size=16)
    idx.add_many(X, labels=lab)
    l, s, _ = idx.topk(q, 25)
    idx.close()
    out["ties_X"], out["ties_lab"], out["ties_q"], out["ties_labels"], out["ties_scores"] = X, lab, q, l, s
    np.savez_compressed(os.path.join(HERE, "flat_cases.npz"), **out)
    print("wrote flat_cases.npz;", len(out), "arrays")


if __name__ == "__main__":
    main()
