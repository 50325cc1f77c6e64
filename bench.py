#!/usr/bin/env python
"""bench.py — flat top-K QPS at batch=1024 (BASELINE.json metric) on N GPUs of one node.

Workload (config.workload): BASELINE.json configs[1] — flat fp32 IP, N=10M d=768 K=100, batch=1024,
synthetic N(0,1) rows L2-normalised. A "step" is one batch of 1024 queries answered
over the whole store. N>1: the store is sharded by contiguous row ranges over the ranks (strong
scaling: total rows fixed), per-shard top-K lists are all-gathered over NCCL and merged on every rank.

  value   whole-job QPS with queries already resident in HBM (vsgpu_topk_device through the C-ABI,
          CUDA-event timed, max over ranks)
  e2e     the same through the reference-facing C API (VecSimIndex_TopKQueryBatchRaw at N=1, the
          sharded front-end at N>1) with HOST query/result buffers: H2D + D2H inside the timed region
  roofline / cpu_baseline   see DESIGN.md §6

  --impl reference   times the reference's own CPU implementation (oracle/_ref, the unmodified
                     sources compiled by oracle/Makefile; falls back to the oracle port) on all host
                     cores, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "flat top-K QPS (batch=1024)"
WORKLOADS = {
    # name: (type, metric, N, dim, K, batch)
    "flat_fp32_ip_10M_d768_k100_b1024": ("fp32", "IP", 10_000_000, 768, 100, 1024),
    "flat_fp32_l2_100k_d128_k10_b1": ("fp32", "L2", 100_000, 128, 10, 1),
    "flat_int8_cos_50M_d512_k10_b4096": ("int8", "Cosine", 50_000_000, 512, 10, 4096),
    "flat_bf16_ip_20M_d1024_k100_b1024": ("bf16", "IP", 20_000_000, 1024, 100, 1024),
    # not a BASELINE config: configs[1] under L2 (the tensor path ranks by a.q - |a|^2 / 2)
    "flat_fp32_l2_10M_d768_k100_b1024": ("fp32", "L2", 10_000_000, 768, 100, 1024),
}
TYPE_ID = {"fp32": 0, "fp64": 1, "bf16": 2, "fp16": 3, "int8": 4, "uint8": 5}
METRIC_ID = {"L2": 0, "IP": 1, "Cosine": 2}
ELEM = {"fp32": 4, "fp64": 8, "bf16": 2, "fp16": 2, "int8": 1, "uint8": 1}
CHUNK = 500_000  # rows generated per chunk (same chunk seeds whatever the sharding)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="flat_fp32_ip_10M_d768_k100_b1024", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override N (debugging only; the line then names it)")
    ap.add_argument("--batch", type=int, default=0, help="override the batch (sweeps; the line then names it)")
    ap.add_argument("--mode", type=int, default=0, help="0 auto, 1 exact scan only, 2 tensor path only")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# synthetic data: chunk c of the global row set depends only on (seed, c)
def gen_chunk_torch(torch, tname, chunk_idx, rows, dim, device, seed=47):
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1_000_003 + chunk_idx)
    if tname == "int8":
        return torch.randint(-128, 128, (rows, dim), generator=g, device=device, dtype=torch.int8)
    x = torch.randn((rows, dim), generator=g, device=device, dtype=torch.float32)
    x = x / x.norm(dim=1, keepdim=True)
    if tname == "bf16":
        return x.to(torch.bfloat16)
    return x


def gen_queries_numpy(tname, nq, dim, seed=48):
    rng = np.random.default_rng(seed)
    if tname == "int8":
        return rng.integers(-128, 128, (nq, dim)).astype(np.int8)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    if tname == "bf16":
        u = q.view(np.uint32)
        return ((u + (((u >> 16) & 1) + 0x7FFF)) >> 16).astype(np.uint16)
    return q


def gen_rows_numpy(tname, n, dim, seed=49):
    rng = np.random.default_rng(seed)
    if tname == "int8":
        return rng.integers(-128, 128, (n, dim)).astype(np.int8)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    if tname == "bf16":
        u = x.view(np.uint32)
        return ((u + (((u >> 16) & 1) + 0x7FFF)) >> 16).astype(np.uint16)
    return x


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel, rows, applicable):
    """DRAM bytes (read + write) of the dominant kernel's launches in one step, from the committed ncu capture
    (profiles/r1_traffic.json: bytes per row measured by ncu x the rows this step streams). None when the capture
    was taken on another shape."""
    if not applicable:
        return None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))[kernel]
        return float(t["bytes_per_row"]) * rows
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return d, "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------
def cpu_reference_qps(tname, mname, n_full, dim, k, batch, seconds, threads=None):
    """Times the reference's CPU path on a bounded sample: n_s rows resident in the reference index,
    queries spread over all host threads (the reference's own pattern: one thread per core pulling
    queries, src/python_bindings/bindings.cpp:250-284). A flat scan is linear in N, so QPS at the full
    N is the sample's QPS * n_s / N — stated in `sample`."""
    threads = threads or os.cpu_count() or 1
    kind = "reference"
    try:
        from oracle import ref
        if not ref.available():
            raise RuntimeError("oracle/_ref not built")
        ref.lib()
    except Exception:
        ref = None
        kind = "port"
    vtype, metric = TYPE_ID[tname], METRIC_ID[mname]
    # size the sample: assume ~8 GB/s/core scanned; aim at `seconds` of wall time
    row_bytes = dim * ELEM[tname]
    n_s = int(min(n_full, max(20_000, 1.2e9 // row_bytes)))     # ~1.2 GB of rows: larger than the host's L3
    X = gen_rows_numpy(tname, n_s, dim)
    Q = gen_queries_numpy(tname, max(threads * 4, 16), dim)
    if ref is not None:
        idx = ref.RefIndex(vtype, dim, metric)
        idx.add_many(X)
        # calibrate then run
        idx.topk_many(Q[:threads], k, n_threads=threads, want_results=False)          # warm-up
        _, _, t1 = idx.topk_many(Q[:threads], k, n_threads=threads, want_results=False)
        per_q = max(t1, 1e-4)  # seconds per round of `threads` queries
        nq = int(max(threads, min(len(Q) * 64, seconds / per_q * threads)))
        Qrun = np.concatenate([Q] * (nq // len(Q) + 1))[:nq]
        _, _, secs = idx.topk_many(Qrun, k, n_threads=threads, want_results=False)
        idx.close()
        used = threads
    else:
        from oracle import port
        idx = port.PortIndex(vtype, dim, metric)
        idx.add_many(X[: min(n_s, 20_000)])
        n_s = idx.size()
        t0 = time.perf_counter()
        nq = 0
        while time.perf_counter() - t0 < min(seconds, 10.0):
            idx.topk(Q[nq % len(Q)], k)
            nq += 1
        secs = time.perf_counter() - t0
        used = 1
    qps_sample = nq / secs
    qps_full = qps_sample * n_s / n_full
    try:
        model = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        model = "unknown"
    return {"value": qps_full, "unit": "queries/s", "cores": used, "kind": kind,
            "sample": "%d queries over %d of %d rows on %d threads in %.1f s (%.1f q/s on the sample), scaled by rows "
                      "(flat scan is linear in N); CPU: %s" % (nq, n_s, n_full, used, secs, qps_sample, model)}


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    tname, mname, n_total, dim, k, batch = WORKLOADS[args.workload]
    if args.rows:
        n_total = args.rows
    if args.batch:
        batch = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl_name = args.workload if not (args.rows or args.batch) else "%s[rows=%d,batch=%d]" % (args.workload, n_total, batch)

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_reference_qps(tname, mname, n_total, dim, k, batch, max(5.0, args.cpu_seconds) * max(1, min(args.steps, 3)) / 3)
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "queries/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": batch / cb["value"] * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32" if tname == "fp32" else tname,
                "data": "synthetic", "config": {"workload": wl_name, "rows": n_total, "dim": dim, "k": k, "batch": batch},
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from vectorsimilarity_b200 import build as vbuild, capi
    if rank == 0:
        vbuild.build()
    if world > 1:
        dist.barrier()
    from vectorsimilarity_b200 import sharded
    capi.lib()
    capi.set_device(local_rank)
    capi.set_topk_mode(args.mode)
    device = torch.device("cuda", local_rank)

    lo, hi = sharded.shard_bounds(n_total, world, rank)
    params = capi.BFParams(type=TYPE_ID[tname], dim=dim, metric=METRIC_ID[mname], multi=False, initialCapacity=hi - lo,
                           blockSize=1024)
    index = sharded.ShardedFlatIndex(params)
    t_ingest = time.perf_counter()
    c0 = lo // CHUNK
    row = lo
    while row < hi:
        c = row // CHUNK
        x = gen_chunk_torch(torch, tname, c, CHUNK, dim, device)
        a, b = row - c * CHUNK, min(hi - c * CHUNK, CHUNK)
        part = x[a:b].contiguous()
        torch.cuda.synchronize()
        index.add_device_rows(part, row)
        row = c * CHUNK + b
        del x, part
    torch.cuda.synchronize()
    t_ingest = time.perf_counter() - t_ingest
    assert index.local.index_size() == hi - lo

    Q = gen_queries_numpy(tname, batch, dim)
    # processed queries for the device-resident leg (IP / L2: the blob itself; cosine would normalise)
    if mname == "Cosine" and tname in ("int8", "uint8"):
        qp = np.zeros((batch, dim + 4), dtype=np.uint8)
        qp[:, :dim] = Q.view(np.uint8)
        norms = np.sqrt((Q.astype(np.int64) ** 2).sum(1).astype(np.float64)).astype(np.float32)
        qp[:, dim:] = norms.view(np.uint8).reshape(batch, 4)
        q_proc = qp
    else:
        q_proc = Q
    q_host = torch.from_numpy(q_proc.view(np.uint8).reshape(batch, -1)).pin_memory()
    q_dev = q_host.to(device)
    torch.cuda.synchronize()

    store_stream = torch.cuda.ExternalStream(sharded._vsgpu().vsgpu_store_stream(index.store()), device=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        return index.topk_device(q_dev, k, args.mode)

    # ---- warm-up ----
    for _ in range(max(args.warmup, 1)):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---- timed: device-resident queries ----
    launches = 0
    scan_ms, total_ms, cands, fallbacks, path = [], [], 0, 0, 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record(store_stream)
    for _ in range(args.steps):
        step_device()
        st = index.last_stats()
        launches += st["kernel_launches"] + (1 if world > 1 else 0)
        cands += st["candidates"]
        fallbacks += st["fallback_queries"]
        path = st["path"]
    # the step's last work is on torch's current stream at N>1 (merge) and on the store stream at N=1
    if world > 1:
        ev1.record(torch.cuda.current_stream())
    else:
        ev1.record(store_stream)
    barrier()
    wall = time.perf_counter() - t_wall0
    dev_ms = ev0.elapsed_time(ev1)
    # per-kernel times of the last step (events recorded inside the C-ABI call on the store's stream)
    st = index.last_stats()
    t = torch.tensor([dev_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = batch / (ms_per_step / 1e3)

    # ---- timed: end to end from host buffers ----
    labels_out = np.empty((batch, k), dtype=np.uint64)
    scores_out = np.empty((batch, k), dtype=np.float64)
    L = capi.lib()

    def step_e2e():
        if world == 1:
            rc = L.VecSimIndex_TopKQueryBatchRaw(index.local._h, Q.ctypes.data, batch, k, None, labels_out.ctypes.data,
                                                 scores_out.ctypes.data)
            assert rc == 0, capi.lib().VecSimGPU_LastError()
        else:
            l, s = index.knn_batch(q_host, k, args.mode)   # pinned host queries in, host labels / scores out
            labels_out[:] = l.view(np.uint64)
            scores_out[:] = s

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_qps = batch * args.steps / e2e_s
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel (last step's CUDA-event time from inside the C-ABI) ----
    peaks, peak_src = measured_peaks()
    n_local = hi - lo
    es = ELEM[tname]
    roof = None
    if rank == 0:
        stl = index.last_stats()
        if stl["path"] == 1 and stl["scan_ms"] > 0 and tname in ("int8", "uint8"):
            # exact integer GEMM (tcgen05 kind::i8) + per-phase merges; no measured int8 peak on this pool: nominal
            # dense int8 is 2x bf16, so the denominator is 2 x the measured sustained bf16 figure (stated)
            ops = 2.0 * batch * n_local * dim
            ach = ops / (stl["scan_ms"] / 1e3) / 1e12
            peak = 2.0 * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s", "frac": ach / peak, "traffic": None,
                    "kernel": "i8_gemm_filter_kernel (+ i8_merge_kernel)", "kernel_ms": stl["scan_ms"],
                    "peak_source": peak_src + " (2 x sustained bf16: nominal int8:bf16 ratio)",
                    "note": "2*B*N*d integer ops over the CUDA-event time of the step's GEMM + merge launches"}
        elif stl["path"] == 1 and stl["scan_ms"] > 0:
            flops = 2.0 * batch * n_local * dim
            ach = flops / (stl["scan_ms"] / 1e3) / 1e12
            peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": ncu_traffic("coarse_gemm_filter_kernel", n_local, tname == "fp32" and dim == 768),
                    "kernel": "coarse_gemm_filter_kernel", "kernel_ms": stl["scan_ms"], "peak_source": peak_src + " (sustained bf16)",
                    "note": "2*B*N*d flops over the summed CUDA-event time of the step's coarse GEMM launches (one per phase)"}
        elif stl["scan_ms"] > 0:
            # exact path: one scan launch streams the shard once for a chunk of <=16 queries
            qc = min(batch, 16)
            byts = n_local * dim * es + qc * dim * es + qc * n_local * 4
            ach = byts / (stl["scan_ms"] / 1e3) / 1e9
            peak = float(peaks.get("hbm_gbs", 6650.0))
            staged = tname in ("fp32", "fp16") and dim % 32 == 0 and not os.environ.get("VSGPU_LEGACY_SCAN")
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": ncu_traffic("scan_tma_kernel_qc16" if qc > 8 else "scan_tma_kernel_qc1", n_local,
                                           staged and tname == "fp32" and dim == 768),
                    "kernel": "scan_tma_kernel" if staged else "exact_scan_kernel", "kernel_ms": stl["scan_ms"], "peak_source": peak_src,
                    "note": "per launch: N*d*s + qc*d*s + qc*N*4 bytes, qc=%d queries per launch" % qc}

    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cpu_base = cpu_reference_qps(tname, mname, n_total, dim, k, batch, args.cpu_seconds)
        except Exception as e:  # the baseline is reported, never required
            cpu_base = {"value": None, "unit": "queries/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if tname == "fp32" else tname, "data": "synthetic",
            "config": {"workload": wl_name, "type": tname, "space": mname, "rows": n_total, "rows_per_gpu": n_local, "dim": dim,
                       "k": k, "batch": batch, "path": ("exact integer GEMM (kind::i8) + merge" if tname in ("int8", "uint8") else
                                "tensor coarse + exact re-rank") if path == 1 else "exact scan",
                       "l2_between_iters": "store (%.1f GB/GPU) is larger than L2" % (n_local * dim * es / 1e9),
                       "sharding": "contiguous row ranges, all-gather of per-shard top-K + merge" if world > 1 else "single GPU",
                       "ingest_s": round(t_ingest, 2)},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(q_proc.nbytes) if world > 1 else int(Q.nbytes),
                    "d2h_bytes_per_step": int(batch * k * (8 + (8 if tname == "fp64" else 4) + 4)) if world == 1 else int(batch * k * 12)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
            "candidates_per_query": cands / max(1, args.steps * batch), "fallback_queries": fallbacks,
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line))
    index.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
