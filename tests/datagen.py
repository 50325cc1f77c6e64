"""Seeded synthetic vectors shared by the parity tests."""
import numpy as np

FLOAT32, FLOAT64, BFLOAT16, FLOAT16, INT8, UINT8 = range(6)
L2, IP, COSINE = range(3)
TYPE_NAMES = ["fp32", "fp64", "bf16", "fp16", "int8", "uint8"]
METRIC_NAMES = ["L2", "IP", "Cosine"]


def to_bf16(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    u = u + (((u >> 16) & 1) + 0x7FFF)
    return (u >> 16).astype(np.uint16)


def from_bf16(h):
    return (np.ascontiguousarray(h, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32)


def make_vectors(vtype, n, dim, seed, dist="uniform"):
    """n x dim raw caller blobs (bf16 as uint16 bit patterns)."""
    rng = np.random.default_rng(seed)
    if vtype in (INT8, UINT8):
        if vtype == INT8:
            return rng.integers(-128, 128, (n, dim)).astype(np.int8)
        return rng.integers(0, 256, (n, dim)).astype(np.uint8)
    if dist == "uniform":
        x = rng.uniform(-1, 1, (n, dim))
    elif dist == "normal":
        x = rng.standard_normal((n, dim))
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    elif dist == "grid":  # small integers: many exact ties
        x = rng.integers(-3, 4, (n, dim)).astype(np.float64)
    else:
        raise ValueError(dist)
    if vtype == FLOAT32:
        return x.astype(np.float32)
    if vtype == FLOAT64:
        return x.astype(np.float64)
    if vtype == FLOAT16:
        return x.astype(np.float16)
    return to_bf16(x.astype(np.float32))
