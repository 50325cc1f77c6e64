"""vectorsimilarity_b200 — B200-native flat top-K / range / HNSW-search hot path of RedisAI/VectorSimilarity.

Product code: csrc/*.cu (sm_100a kernels + C-ABI, libvsgpu.so), csrc/host/*.cpp (VecSimIndex_* C API,
libvecsim_b200.so), capi.py (ctypes mirror of the reference's Python binding), sharded.py (one process
per GPU, NCCL all-gather of per-shard top-K). Nothing here imports oracle/.
"""
from . import capi  # noqa: F401
from .capi import (BFIndex, BFParams, BY_ID, BY_SCORE, HNSWIndex, HNSWParams, VecSimMetric_Cosine, VecSimMetric_IP, VecSimMetric_L2,  # noqa: F401
                   VecSimType_BFLOAT16, VecSimType_FLOAT16, VecSimType_FLOAT32, VecSimType_FLOAT64, VecSimType_INT8,
                   VecSimType_UINT8)
