"""Golden case for the serialized-HNSW-index reader.

tests/golden/ref_hnsw_1k_d4_single.v3 is the reference's own test fixture
(/root/reference/tests/unit/data/1k-d4-L2-M8-ef_c10_FLOAT32_single.v3, loaded by tests/unit/test_hnsw.cpp:1991-2059):
an encoding-V3 file of 1001 fp32 vectors (dim 4, L2, M 8, efConstruction 10, blockSize 2) written by the reference's
HNSWIndex::saveIndex. The reference's loader is compiled only under BUILD_TESTS, which oracle/_ref does not define, so the
expected answers are produced by the UNMODIFIED reference index (oracle/_ref) rebuilt from the file's vectors in id order
— after checking that this rebuild reproduces the file's graph link for link (same levels, same link lists, same entry
point), i.e. that the answers are the ones the reference gives on the file's own graph.

    python tests/golden/make_hnsw_file_golden.py
"""
import hashlib
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import port, ref  # noqa: E402

SRC = "/root/reference/tests/unit/data/1k-d4-L2-M8-ef_c10_FLOAT32_single.v3"
DST = os.path.join(HERE, "ref_hnsw_1k_d4_single.v3")
NQ, K = 16, 10


def main():
    shutil.copyfile(SRC, DST)
    f = port.read_hnsw_file(DST)
    X = np.ascontiguousarray(f["vectors"]).view(np.float32)
    ref.lib()
    R = ref.RefIndex(0, f["dim"], f["metric"], algo="hnsw", M=f["M"], ef_construction=f["ef_construction"],
                     ef_runtime=f["ef_runtime"])
    R.add_many(X, labels=f["labels"])
    g = R.hnsw_export()
    assert g["entry"] == f["entry"] and g["max_level"] == f["max_level"] and np.array_equal(g["levels"], f["levels"])
    for lvl in range(len(f["links"])):
        assert np.array_equal(g["counts"][lvl], f["counts"][lvl]), lvl
        w = f["links"][lvl].shape[1]
        mask = np.arange(w)[None, :] < f["counts"][lvl][:, None]
        assert np.array_equal(np.where(mask, g["links"][lvl], 0), np.where(mask, f["links"][lvl], 0)), lvl
    rng = np.random.default_rng(99)
    Q = rng.uniform(0, 1, (NQ, f["dim"])).astype(np.float32)
    out = dict(Q=Q, sha256=np.frombuffer(hashlib.sha256(open(DST, "rb").read()).digest(), dtype=np.uint8))
    for ef in (10, 50):
        labels = np.zeros((NQ, K), dtype=np.int64)
        scores = np.zeros((NQ, K))
        for i in range(NQ):
            l, s, _ = R.topk(Q[i], K, ef_runtime=ef)
            labels[i], scores[i] = l, s
        out[f"labels_ef{ef}"], out[f"scores_ef{ef}"] = labels, scores
    rl, rs, _ = R.range(Q[0], 0.05)
    out["range_labels"], out["range_scores"] = rl.astype(np.int64), rs
    np.savez_compressed(os.path.join(HERE, "hnsw_file_case.npz"), **out)
    print("wrote", DST, "and hnsw_file_case.npz;", len(rl), "range results")


if __name__ == "__main__":
    main()
