"""CPU-only checks of the product's host side: the C-ABI libraries load and export every symbol the
headers declare, fail loudly without a device, and the tie-resolution logic that
turns the device's (score, id)-ordered candidates into the reference's reply matches the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def capi():
    from vectorsimilarity_b200 import build, capi as c
    build.build()
    c.lib()
    return c


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:VecSim|vsgpu_)[A-Za-z0-9_]*)\s*\(", src)) - {"VecSim_OK"})


def test_vecsim_header_symbols_exported(capi):
    L = capi.lib()
    names = _declared("vecsim_b200.h")
    assert len(names) > 50
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == sorted(names)


def test_vsgpu_header_symbols_exported(capi):
    G = C.CDLL(os.path.join(ROOT, "vectorsimilarity_b200", "libvsgpu.so"))
    names = _declared("vsgpu.h")
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(G, n)]
    assert not missing, missing


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    p = capi.BFParams(type=0, dim=4, metric=0, multi=False, initialCapacity=0, blockSize=0)
    with pytest.raises(RuntimeError, match="no such CUDA device|NULL"):
        capi.BFIndex(p)


def test_product_does_not_touch_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "vectorsimilarity_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "vs_oracle" not in text and "libvecsim_ref" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), f


def test_blob_sizes_and_normalize_match_oracle(capi, port):
    L = capi.lib()
    rng = np.random.default_rng(3)
    for vtype in range(6):
        for metric in range(3):
            for dim in (1, 7, 128, 513):
                assert L.VecSimParams_GetQueryBlobSize(vtype, dim, metric) == port.stored_size(vtype, metric, dim)
    from datagen import make_vectors
    for vtype in range(6):
        for dim in (3, 16, 100):
            raw = make_vectors(vtype, 2, dim, seed=dim)
            for i in range(2):
                if vtype in (4, 5):
                    a = np.zeros(dim + 4, dtype=np.uint8)
                    a[:dim] = raw[i].view(np.uint8)
                else:
                    a = raw[i].copy()
                b = a.copy()
                capi.normalize(a, dim, vtype)
                port.normalize(vtype, dim, b)
                assert a.tobytes() == b.tobytes()


def test_tie_resolution_matches_oracle(capi, port):
    """Feed the host resolver what the device would return — the min(2k, n) best rows by
    (score, internal id) — and compare with the oracle's sequential heap scan."""
    L = capi.lib()
    L.vsb_test_resolve.restype = C.c_size_t
    L.vsb_test_resolve.argtypes = [C.c_void_p] * 3 + [C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(11)
    for trial in range(300):
        n, dim, k = int(rng.integers(5, 150)), int(rng.integers(1, 4)), int(rng.integers(1, 40))
        X = rng.integers(-2, 3, (n, dim)).astype(np.int8)
        labels = rng.permutation(n * 4)[:n].astype(np.uint64)
        q = rng.integers(-2, 3, dim).astype(np.int8)
        P = port.PortIndex(4, dim, 0)
        P.add_many(X, labels=labels)
        want_l, want_s, _ = P.topk(q, k)
        P.close()
        scores = ((X.astype(np.int64) - q.astype(np.int64)) ** 2).sum(1).astype(np.float64)
        order = np.lexsort((np.arange(n), scores))          # ascending (score, id)
        k_sel = n if k > n // 2 else min(2 * k, n)
        sel = order[:k_sel]
        cl = np.ascontiguousarray(labels[sel])
        cs = np.ascontiguousarray(scores[sel])
        ci = np.ascontiguousarray(sel.astype(np.uint32))
        out_l = np.empty(max(k, 1), dtype=np.uint64)
        out_s = np.empty(max(k, 1), dtype=np.float64)
        m = L.vsb_test_resolve(cl.ctypes.data, cs.ctypes.data, ci.ctypes.data, k_sel, k, out_l.ctypes.data,
                               out_s.ctypes.data)
        assert m == len(want_l), trial
        assert np.array_equal(out_l[:m], want_l), trial
        assert np.array_equal(out_s[:m], want_s), trial


def test_tiered_merge_matches_reference_merge_results(capi, port):
    """The tiered index's reply merge (csrc/host/vecsim_tiered.cpp) against the restated merge_results
    (query_result_utils.h:44-92): score-then-id order with the 1e-6 score epsilon, duplicates emitted once, limit."""
    L = capi.lib()
    L.vsb_test_tiered_merge.restype = C.c_size_t
    L.vsb_test_tiered_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(5)
    for trial in range(400):
        n_all = int(rng.integers(0, 60))
        ids = rng.permutation(200)[:n_all].astype(np.uint64)
        # few distinct scores (ties), some within the epsilon of each other
        scores = rng.integers(0, 8, n_all).astype(np.float64) + rng.choice([0.0, 2e-7, 4e-7], n_all)
        in_a = rng.random(n_all) < 0.6
        in_b = (rng.random(n_all) < 0.6) | ~in_a
        def side(mask):
            sel = np.nonzero(mask)[0]
            order = np.lexsort((ids[sel], scores[sel]))
            return np.ascontiguousarray(ids[sel][order]), np.ascontiguousarray(scores[sel][order])
        ai, asc = side(in_a)
        bi, bsc = side(in_b)
        limit = int(rng.integers(0, 70)) if trial % 5 else (1 << 64) - 1
        want, ta, tb = port.merge_results(list(zip(ai.tolist(), asc.tolist())), list(zip(bi.tolist(), bsc.tolist())),
                                          None if limit == (1 << 64) - 1 else limit)
        out_i = np.empty(len(ai) + len(bi) + 1, dtype=np.uint64)
        out_s = np.empty(len(ai) + len(bi) + 1, dtype=np.float64)
        taken = np.zeros(2, dtype=np.uint64)
        m = L.vsb_test_tiered_merge(ai.ctypes.data, asc.ctypes.data, len(ai), bi.ctypes.data, bsc.ctypes.data, len(bi), limit,
                                    out_i.ctypes.data, out_s.ctypes.data, taken.ctypes.data)
        assert m == len(want), trial
        assert out_i[:m].tolist() == [w[0] for w in want], trial
        assert out_s[:m].tolist() == [w[1] for w in want], trial
        assert (int(taken[0]), int(taken[1])) == (ta, tb), trial


def test_hybrid_policy_cost_model(capi):
    """preferAdHocSearch is a cost comparison with device rates (csrc/host/vecsim_hybrid.h): ad hoc
    for an empty index; once batches win at some subset size they win for every larger one; a flat index prefers ad hoc
    until the per-label host work outweighs one streamed pass plus its score transfer and selections; on HNSW the
    crossover grows like sqrt(k * N)."""
    L = capi.lib()
    L.vsb_test_prefer_adhoc.argtypes = [C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
    for algo in (0, 1):
        assert L.vsb_test_prefer_adhoc(algo, 0, 512, 0, 10) == 1
        for n, rb, k in ((10_000, 512, 10), (1_000_000, 3072, 100), (50_000_000, 516, 10)):
            prev = 1
            for frac in (0.0001, 0.001, 0.01, 0.05, 0.2, 0.5, 0.9, 1.0, 5.0):
                cur = L.vsb_test_prefer_adhoc(algo, n, rb, max(1, int(frac * n)), k)
                assert not (cur == 1 and prev == 0), (algo, n, frac)      # monotone: never back to ad hoc
                prev = cur
    # flat, 10 M rows of 3 KB: a batch pass costs ~48 ms (scan 4.7 + scores to the host 3.2 + two host selections 40), an
    # ad-hoc label ~31 ns (gather 1 ns, label lookup 30 ns): crossover near 1.5 M labels
    assert L.vsb_test_prefer_adhoc(0, 10_000_000, 3072, 500_000, 100) == 1
    assert L.vsb_test_prefer_adhoc(0, 10_000_000, 3072, 5_000_000, 100) == 0
    # HNSW, 1 M rows of 512 B, k = 10: crossover near sqrt(hop * k * N / per-label cost) ~ 50 K labels
    assert L.vsb_test_prefer_adhoc(1, 1_000_000, 512, 10_000, 10) == 1
    assert L.vsb_test_prefer_adhoc(1, 1_000_000, 512, 200_000, 10) == 0
    assert L.vsb_test_prefer_adhoc(1, 1_000_000, 512, 200_000, 1000) == 1        # many results wanted: batches get long


def test_header_is_c_and_a_c_consumer_links(capi, tmp_path):
    """include/vecsim_b200.h is what a C consumer (RediSearch) includes: a C11 program that fills the parameter structs
    with designated initialisers — including the tiered block with a submit callback — compiles without warnings, links
    against libvecsim_b200.so and sees the reference's struct sizes; without a device VecSimIndex_New fails loudly."""
    import subprocess
    exe = str(tmp_path / "consumer")
    libdir = os.path.join(ROOT, "vectorsimilarity_b200")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c", "consumer.c"), "-o", exe, "-L" + libdir, "-lvecsim_b200",
                        "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0] == "136 64 360"
    assert lines[1].startswith("index ") and ("(ok)" in lines[1] or "tiered index:" in lines[1])


def test_shard_bound_exchange_is_a_lower_bound_of_the_global_kth():
    """The arithmetic behind vsgpu_topk_device_begin / _next / _finish, on a numpy model: every row has an
    exact score inside [lo, hi]; a shard reports the k-th and the ceil(k / W)-th largest `lo` of the rows it has seen (or -inf);
    B = max(max_s kth_s, min_s mth_s) must never exceed the k-th largest exact score over all shards, so every row of the global
    top-k — ties at the k-th score included — passes the admission test hi >= B. Also with shards of very different sizes,
    shards that have seen fewer than m rows, heavy ties and zero-width intervals."""
    rng = np.random.default_rng(7)
    for trial in range(300):
        W = int(rng.integers(2, 9))
        k = int(rng.integers(1, 40))
        m = -(-k // W)
        sizes = rng.integers(0, 60, size=W)
        if trial % 3 == 0:
            sizes[rng.integers(0, W)] = 200                       # one shard holds most of the rows
        if sizes.sum() < k:
            sizes[0] += k
        exact, lo, hi = [], [], []
        for n in sizes:
            e = rng.normal(size=n)
            if trial % 4 == 1:
                e = np.round(e * 2) / 2                            # many exact ties
            w = rng.uniform(0, 0.3, size=n) if trial % 5 else np.zeros(n)
            shift = rng.uniform(-1, 1, size=n)
            exact.append(e)
            lo.append(e - w * (1 + shift) / 2 - 0.0)
            hi.append(e + w * (1 - shift) / 2 + 0.0)
        def nth_largest(a, r):
            return -np.inf if len(a) < r else np.sort(a)[::-1][r - 1]
        kth = max(nth_largest(l, k) for l in lo)
        mth = min(nth_largest(l, m) for l in lo)
        B = max(kth, mth)
        all_exact = np.concatenate(exact)
        true_kth = np.sort(all_exact)[::-1][k - 1]
        assert B <= true_kth + 1e-12, (trial, B, true_kth)
        all_hi = np.concatenate(hi)
        in_topk = all_exact >= true_kth                             # the top-k and everything tied with its last member
        assert np.all(all_hi[in_topk] >= B)


def test_phase_schedule_of_sharded_phased_calls(capi):
    """make_phases with `world` shards and a fixed number of rounds (vsgpu_topk_device_begin / _next / _finish): every shard
    runs exactly `rounds` phases whatever its own row count (trailing ones may be empty), the schedule still tiles [0, n), and
    exchanging the ceil(k / world)-th best affords fewer phases than a shard on its own needs."""
    capi.lib()
    G = C.CDLL(os.path.join(ROOT, "vectorsimilarity_b200", "libvsgpu.so"))
    G.vsgpu_debug_phases.restype = C.c_size_t
    G.vsgpu_debug_phases.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
    G.vsgpu_debug_phases_sharded.restype = C.c_size_t
    G.vsgpu_debug_phases_sharded.argtypes = [C.c_size_t, C.c_size_t, C.c_uint, C.c_size_t, C.c_void_p, C.c_size_t]
    G.vsgpu_topk_rounds.restype = C.c_size_t
    G.vsgpu_topk_rounds.argtypes = [C.c_size_t, C.c_size_t, C.c_uint]
    buf = np.zeros(64, dtype=np.uint32)
    for n, k, world in ((1_250_000, 100, 8), (2_500_000, 100, 4), (5_000_000, 100, 2), (6_250_000, 10, 8), (40_000, 384, 2),
                        (1_250_000, 1000, 8)):
        rounds = G.vsgpu_topk_rounds(n, k, world)
        own = G.vsgpu_debug_phases(n, k, buf.ctypes.data, 64)
        assert 1 <= rounds <= own
        if world >= 4 and k <= 100:
            assert rounds < own
        for rows in (n, n - 1, n // 3, 1000, 128, 77):     # shards of any size follow the agreed number of rounds
            m = G.vsgpu_debug_phases_sharded(rows, k, world, rounds, buf.ctypes.data, 64)
            edges = buf[:m].astype(np.int64)
            assert m == rounds and edges[-1] == rows and np.all(np.diff(edges) >= 0)
            assert np.all(edges[edges < rows] % 128 == 0)
    assert G.vsgpu_topk_rounds(1_250_000, 100, 8) == 3
    assert G.vsgpu_topk_rounds(10_000_000, 100, 1) == G.vsgpu_debug_phases(10_000_000, 100, buf.ctypes.data, 64)


def test_phase_schedule_of_the_filtered_gemm(capi):
    """make_phases (csrc/vsgpu_tc.cuh): phases tile [0, n) without gaps, every edge but the last is a multiple of the
    128-row tile, consecutive edges grow by at most the factor the candidate buffer affords (8 for k <= 150, less for
    larger k), and no schedule with one phase fewer would satisfy that."""
    capi.lib()
    G = C.CDLL(os.path.join(ROOT, "vectorsimilarity_b200", "libvsgpu.so"))
    G.vsgpu_debug_phases.restype = C.c_size_t
    G.vsgpu_debug_phases.argtypes = [C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
    buf = np.zeros(64, dtype=np.uint32)
    for n, k in ((32768, 10), (100_000, 1), (1_250_000, 100), (10_000_000, 100), (20_000_000, 100), (6_250_000, 10),
                 (40_000, 384), (50_000_000, 10), (3_000_001, 37)):
        m = G.vsgpu_debug_phases(n, k, buf.ctypes.data, 64)
        edges = buf[:m].astype(np.int64)
        assert m >= 1 and edges[-1] == n and np.all(np.diff(edges) > 0)
        assert np.all(edges[:-1] % 128 == 0)
        growth = max(3.0, min(8.0, 3072 / (2.5 * k)))
        s0 = edges[0]
        assert s0 >= min(n, 2048) and s0 >= min(n, 2 * k)
        ratios = edges[1:] / edges[:-1]
        assert np.all(ratios <= growth * 1.02), (n, k, ratios)
        if m > 2:  # fewest phases: one fewer would need a larger ratio than allowed
            assert (n / s0) ** (1.0 / (m - 2)) > growth * 0.999, (n, k, m)
