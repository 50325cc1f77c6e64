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
    ap.add_argument("--extras", default="auto", choices=["auto", "none"],
                    help="auto: batch sweep + the other BASELINE configs as sub-results of the line; none: headline only")
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
class FlatRun:
    """One flat workload resident on this rank's GPU: ingest, then device-timed and end-to-end steps."""

    def __init__(self, env, wl_name, rows_total, batch):
        self.env = env
        torch, capi, sharded = env["torch"], env["capi"], env["sharded"]
        self.tname, self.mname, _, self.dim, self.k, b0 = WORKLOADS[wl_name]
        self.batch = batch or b0
        self.rows_total = rows_total
        self.lo, self.hi = sharded.shard_bounds(rows_total, env["world"], env["rank"])
        self.n_local = self.hi - self.lo
        params = capi.BFParams(type=TYPE_ID[self.tname], dim=self.dim, metric=METRIC_ID[self.mname], multi=False,
                               initialCapacity=self.n_local, blockSize=1024)
        self.index = sharded.ShardedFlatIndex(params)
        t0 = time.perf_counter()
        row = self.lo
        while row < self.hi:
            c = row // CHUNK
            x = gen_chunk_torch(torch, self.tname, c, CHUNK, self.dim, env["device"])
            a, b = row - c * CHUNK, min(self.hi - c * CHUNK, CHUNK)
            part = x[a:b].contiguous()
            torch.cuda.synchronize()
            self.index.add_device_rows(part, row)
            row = c * CHUNK + b
            del x, part
        torch.cuda.synchronize()
        self.ingest_s = time.perf_counter() - t0
        assert self.index.local.index_size() == self.n_local
        self.set_batch(self.batch)

    def set_batch(self, batch):
        torch = self.env["torch"]
        self.batch = batch
        dim, tname, mname = self.dim, self.tname, self.mname
        self.Q = gen_queries_numpy(tname, batch, dim)
        # processed queries for the device-resident leg (IP / L2: the blob itself; int8 cosine appends the norm)
        if mname == "Cosine" and tname in ("int8", "uint8"):
            qp = np.zeros((batch, dim + 4), dtype=np.uint8)
            qp[:, :dim] = self.Q.view(np.uint8)
            norms = np.sqrt((self.Q.astype(np.int64) ** 2).sum(1).astype(np.float64)).astype(np.float32)
            qp[:, dim:] = norms.view(np.uint8).reshape(batch, 4)
            self.q_proc = qp
        else:
            self.q_proc = self.Q
        self.q_host = torch.from_numpy(self.q_proc.view(np.uint8).reshape(batch, -1)).pin_memory()
        self.q_dev = self.q_host.to(self.env["device"])
        torch.cuda.synchronize()

    def barrier(self):
        if self.env["world"] > 1:
            self.env["dist"].barrier()
        self.env["torch"].cuda.synchronize()

    def time_device(self, steps, warmup, mode):
        """W warm-up steps, then exactly `steps` steps with device-resident queries: CUDA events on the store's stream
        (every step — scan, collective, merge — is enqueued there), max over ranks."""
        torch, world = self.env["torch"], self.env["world"]
        index, st = self.index, self.index.stream()
        for _ in range(max(warmup, 1)):
            index.topk_device(self.q_dev, self.k, mode)
            index.finish()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches, cands, fallbacks, path, scan_ms, local_ms = 0, 0, 0, 0, [], []
        self.barrier()
        t0 = time.perf_counter()
        ev0.record(st)
        for _ in range(steps):
            index.topk_device(self.q_dev, self.k, mode)
            index.finish()      # end of the step: the result is final (overflowed queries redone), one host wait per step
            s = index.last_stats()
            launches += s["kernel_launches"] + (2 if world > 1 else 0)   # + pack / merge of the gathered hits
            cands += s["candidates"]
            fallbacks += s["fallback_queries"]
            path = s["path"]
            scan_ms.append(s["scan_ms"])
            local_ms.append(s["total_ms"])
        ev1.record(st)
        self.barrier()
        wall = time.perf_counter() - t0
        t = torch.tensor([ev0.elapsed_time(ev1)], device=self.env["device"], dtype=torch.float64)
        if world > 1:
            self.env["dist"].all_reduce(t, op=self.env["dist"].ReduceOp.MAX)
        ms = float(t.item()) / steps
        return {"ms_per_step": ms, "qps": self.batch / (ms / 1e3), "launches": launches, "candidates": cands,
                "fallbacks": fallbacks, "path": path, "scan_ms": float(np.mean(scan_ms)) if scan_ms else 0.0,
                "local_ms": float(np.mean(local_ms)) if local_ms else 0.0, "wall": wall}

    def time_e2e(self, steps, warmup, mode):
        """The same steps through the reference-facing call with HOST buffers (H2D of the queries and D2H of the reply
        inside the timed region): VecSimIndex_TopKQueryBatchRaw at N=1, the sharded front-end at N>1."""
        torch, capi, world = self.env["torch"], self.env["capi"], self.env["world"]
        batch, k = self.batch, self.k
        labels_out = np.empty((batch, k), dtype=np.uint64)
        scores_out = np.empty((batch, k), dtype=np.float64)
        L = capi.lib()

        def step():
            if world == 1:
                rc = L.VecSimIndex_TopKQueryBatchRaw(self.index.local._h, self.Q.ctypes.data, batch, k, None,
                                                     labels_out.ctypes.data, scores_out.ctypes.data)
                assert rc == 0, L.VecSimGPU_LastError()
            else:
                # pinned host queries in, host labels / scores out
                self.index.knn_batch(self.q_host, k, mode, out_labels=labels_out, out_scores=scores_out)

        for _ in range(max(1, min(warmup, 2))):
            step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.barrier()
        t = torch.tensor([time.perf_counter() - t0], device=self.env["device"], dtype=torch.float64)
        if world > 1:
            self.env["dist"].all_reduce(t, op=self.env["dist"].ReduceOp.MAX)
        secs = float(t.item())
        h2d = int(self.q_proc.nbytes) if world > 1 else int(self.Q.nbytes)
        d2h = int(batch * k * (8 + 4 + 4)) if world == 1 else int(batch * k * 12)
        return {"value": batch * steps / secs, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}

    def bounds(self, ms, peaks):
        """algorithmic bytes and ops of one batch over this rank's shard against both rooflines."""
        es = ELEM[self.tname]
        norms = 4 * self.n_local if (self.tname in ("int8", "uint8") and self.mname == "Cosine") or self.mname == "L2" else 0
        byts = self.n_local * self.dim * es + norms + self.batch * self.dim * es + self.batch * self.k * 12
        ops = 2.0 * self.batch * self.n_local * self.dim
        hbm = float(peaks.get("hbm_gbs", 6650.0)) * 1e9
        tens = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0))) * 1e12
        if self.tname in ("int8", "uint8"):
            tens *= float(peaks.get("int8_vs_bf16", 2.0))
        t = ms / 1e3
        t_bound = max(byts / hbm, ops / tens)
        return {"hbm_fraction": byts / t / hbm, "tensor_fraction": ops / t / tens, "min_bound_fraction": t_bound / t,
                "bound": "hbm" if byts / hbm >= ops / tens else "tensor", "algorithmic_bytes": byts, "ops": ops}

    def roofline(self, res, peaks, peak_src):
        """roofline object of the dominant kernel, from the CUDA-event time recorded inside the C-ABI call."""
        if res["scan_ms"] <= 0:
            return None
        n_local, batch, dim, tname = self.n_local, self.batch, self.dim, self.tname
        es = ELEM[tname]
        if res["path"] == 1 and tname in ("int8", "uint8"):
            ops = 2.0 * batch * n_local * dim
            ach = ops / (res["scan_ms"] / 1e3) / 1e12
            ratio = float(peaks.get("int8_vs_bf16", 2.0))
            peak = ratio * float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
            return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TOP/s", "frac": ach / peak, "traffic": None,
                    "kernel": "i8_gemm_filter_kernel (+ i8_merge_kernel)", "kernel_ms": res["scan_ms"],
                    "peak_source": peak_src + " (%.2f x sustained bf16: %s)" % (ratio, peaks.get("int8_source", "nominal int8:bf16 ratio")),
                    "note": "2*B*N*d integer ops over the CUDA-event time of the step's GEMM + merge launches"}
        if res["path"] == 1:
            flops = 2.0 * batch * n_local * dim
            ach = flops / (res["scan_ms"] / 1e3) / 1e12
            peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
            return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": ncu_traffic("coarse_gemm_filter_kernel", n_local, tname == "fp32" and dim == 768),
                    "kernel": "coarse_gemm_filter_kernel", "kernel_ms": res["scan_ms"], "peak_source": peak_src + " (sustained bf16)",
                    "note": "2*B*N*d flops over the summed CUDA-event time of the step's coarse GEMM launches (one per phase)"}
        # exact path: one scan launch streams the shard once for a chunk of <= 16 queries
        qc = min(batch, 16)
        byts = n_local * dim * es + qc * dim * es + qc * n_local * 4
        ach = byts / (res["scan_ms"] / 1e3) / 1e9
        peak = float(peaks.get("hbm_gbs", 6650.0))
        staged = tname in ("fp32", "fp16") and dim % 32 == 0 and not os.environ.get("VSGPU_LEGACY_SCAN")
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": ncu_traffic("scan_tma_kernel_qc16" if qc > 8 else "scan_tma_kernel_qc1", n_local,
                                       staged and tname == "fp32" and dim == 768),
                "kernel": "scan_tma_kernel" if staged else "exact_scan_kernel", "kernel_ms": res["scan_ms"], "peak_source": peak_src,
                "note": "per launch: N*d*s + qc*d*s + qc*N*4 bytes, qc=%d queries per launch (last of the step's launches)" % qc}

    def path_name(self, path):
        if path != 1:
            return "exact scan"
        return "exact integer GEMM (kind::i8) + merge" if self.tname in ("int8", "uint8") else "tensor coarse + exact re-rank"

    def close(self):
        self.index.close()
        self.env["torch"].cuda.empty_cache()


def sub_config(env, wl_name, rows_total, steps, warmup, mode, peaks, e2e=True):
    """A further BASELINE config as a sub-result of the line (same measurement as the headline workload)."""
    r = FlatRun(env, wl_name, rows_total, 0)
    res = r.time_device(steps, warmup, mode)
    out = None
    e = r.time_e2e(steps, warmup, mode) if e2e else None
    if env["rank"] == 0:
        roof = r.roofline(res, peaks, env["peak_src"])
        out = {"workload": wl_name, "rows": rows_total, "rows_per_gpu": r.n_local, "dim": r.dim, "k": r.k, "batch": r.batch,
               "value": res["qps"], "unit": "queries/s", "ms_per_step": res["ms_per_step"], "e2e": e, "path": r.path_name(res["path"]),
               "gpu_launches": res["launches"], "fallback_queries": res["fallbacks"],
               "candidates_per_query": res["candidates"] / max(1, steps * r.batch), "roofline": roof,
               "ingest_s": round(r.ingest_s, 2)}
        out.update(r.bounds(res["ms_per_step"], peaks))
    r.close()
    return out


def single_process_sharded(env, wl_name, n_total, steps, warmup, mode):
    """The same workload through the sharding INSIDE the C library (VecSimGPU_Configure + VecSimIndex_New(VecSimAlgo_BF),
    one process driving every GPU of the box; csrc/host/vecsim_flat_sharded.cpp): what RediSearch gets by linking
    libvecsim_b200.so. Runs on rank 0 only, after the SPMD measurement, with HOST query / result buffers
    (VecSimIndex_TopKQueryBatchRaw: H2D to every device, per-shard scan, peer-copy gather, merge on device 0, D2H)."""
    torch, capi = env["torch"], env["capi"]
    tname, mname, _, dim, k, batch = WORKLOADS[wl_name]
    ndev = env["world"]
    capi.configure_devices(list(range(ndev)))
    capi.set_topk_mode(mode)
    try:
        G = capi.BFIndex(capi.BFParams(type=TYPE_ID[tname], dim=dim, metric=METRIC_ID[mname], multi=False, initialCapacity=n_total,
                                       blockSize=1024))
        assert G.shard_count() == ndev
        t0 = time.perf_counter()
        for d in range(ndev):
            lo, hi = env["sharded"].shard_bounds(n_total, ndev, d)
            dev = torch.device("cuda", d)
            row = lo
            while row < hi:
                c = row // CHUNK
                x = gen_chunk_torch(torch, tname, c, CHUNK, dim, dev)
                a, b = row - c * CHUNK, min(hi - c * CHUNK, CHUNK)
                part = x[a:b].contiguous()
                torch.cuda.synchronize(dev)
                assert G.add_device_rows(part.data_ptr(), part.stride(0) * part.element_size(), b - a, row) == b - a
                row = c * CHUNK + b
                del x, part
        ingest_s = time.perf_counter() - t0
        assert G.index_size() == n_total
        Q = gen_queries_numpy(tname, batch, dim)
        labels_out = np.empty((batch, k), dtype=np.uint64)
        scores_out = np.empty((batch, k), dtype=np.float64)
        L = capi.lib()

        def step():
            rc = L.VecSimIndex_TopKQueryBatchRaw(G._h, Q.ctypes.data, batch, k, None, labels_out.ctypes.data, scores_out.ctypes.data)
            assert rc == 0, L.VecSimGPU_LastError()

        for _ in range(max(warmup, 1)):
            step()
        dev_ms = []
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
            dev_ms.append(G.last_query_stats()["total_ms"])
        secs = time.perf_counter() - t0
        st = G.last_query_stats()
        out = {"devices": ndev, "rows": n_total, "batch": batch, "k": k,
               "e2e": {"value": batch * steps / secs, "unit": "queries/s", "h2d_bytes_per_step": int(Q.nbytes) * ndev,
                       "d2h_bytes_per_step": int(batch * k * 12)},
               "device_ms_per_step": float(np.mean(dev_ms)), "value": batch / (float(np.mean(dev_ms)) / 1e3) if np.mean(dev_ms) > 0 else None,
               "path": "tensor coarse + exact re-rank" if st["path"] == 1 else "exact scan", "gpu_launches": st["kernel_launches"] * steps,
               "fallback_queries": st["fallback_queries"], "ingest_s": round(ingest_s, 2),
               "how": "one process, one host thread + stream per GPU, peer-copy gather of packed per-shard top-K, merge on device 0"}
        G.close()
        return out
    finally:
        capi.set_device(env["device"].index)


def cfg1_latency(env):
    """BASELINE configs[0] (flat fp32 L2, N=100k d=128 K=10, ONE query): the call RediSearch makes today, through
    VecSimIndex_TopKQuery with a host blob in and a reply object out. Latency-bound: reported as microseconds."""
    capi = env["capi"]
    tname, mname, n, dim, k, _ = WORKLOADS["flat_fp32_l2_100k_d128_k10_b1"]
    rng = np.random.default_rng(47)
    X = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    Q = rng.uniform(-1, 1, (64, dim)).astype(np.float32)
    G = capi.BFIndex(capi.BFParams(type=0, dim=dim, metric=0, multi=False, initialCapacity=n, blockSize=1024))
    G.add_vectors(X)
    for i in range(8):
        G.knn_query(Q[i], k)
    dev, api = [], []
    for i in range(64):
        t0 = time.perf_counter()
        G.knn_query(Q[i], k)
        api.append(time.perf_counter() - t0)
        dev.append(G.last_query_stats()["total_ms"])
    out = {"workload": "flat_fp32_l2_100k_d128_k10_b1", "latency_us_api_median": float(np.median(api)) * 1e6,
           "latency_us_device_median": float(np.median(dev)) * 1e3, "gpu_launches_per_query": G.last_query_stats()["kernel_launches"],
           "qps_single_stream": 1.0 / float(np.median(api))}
    G.close()
    return out


def cfg5_hnsw(env, steps, warmup):
    """BASELINE configs[4]: HNSW fp32 L2, N=1M d=128 M=16 efC=200 efR=64 K=10, batch=256 — the graph the unmodified
    reference built (hnsw_cache/cfg5_graph_1000000.npz: links + the reference's recorded answers; vectors are regenerated
    from the seed). Searched on the device; ids and scores are compared with the recorded answers inside the run."""
    capi = env["capi"]
    n, dim, M, efc, ef, k, nq = 1_000_000, 128, 16, 200, 64, 10, 256
    path = os.path.join(ROOT, "hnsw_cache", "cfg5_graph_%d.npz" % n)
    if not os.path.exists(path):
        return {"workload": "hnsw_fp32_l2_1M_d128_M16_efc200_ef64_k10_b256", "unavailable": "hnsw_cache/ graph file not shipped"}
    g = np.load(path)
    rng = np.random.default_rng(47)
    X = rng.uniform(-1, 1, (n, dim)).astype(np.float32)
    Q = rng.uniform(-1, 1, (nq, dim)).astype(np.float32)
    G = capi.HNSWIndex(capi.HNSWParams(type=0, dim=dim, metric=0, multi=False, initialCapacity=n, blockSize=1024, M=M,
                                       efConstruction=efc, efRuntime=ef, epsilon=0.01))
    levels, l0, upper = (np.ascontiguousarray(g[key]) for key in ("levels", "l0", "upper"))
    t0 = time.perf_counter()
    rc = capi.lib().VecSimGPU_HNSWImportGraph(G._h, X.ctypes.data, 1, n, None, levels.ctypes.data, l0.ctypes.data,
                                              upper.ctypes.data if len(upper) else None, len(upper), int(g["entry"][0]),
                                              int(g["entry"][1]))
    assert rc == 0, capi.lib().VecSimGPU_LastError()
    load_s = time.perf_counter() - t0
    for _ in range(max(warmup, 1)):
        labels, scores = G.knn_batch(Q, k)
    ms, wall = [], []
    for _ in range(steps):
        t0 = time.perf_counter()
        labels, scores = G.knn_batch(Q, k)       # host queries in, host labels / scores out
        wall.append(time.perf_counter() - t0)
        ms.append(G.hnsw_stats()["ms"])
    st = G.hnsw_stats()
    kms = float(np.mean(ms))
    evals, hops = st["dist_evals"], st["hops"]
    touched = evals * (dim * 4 + 4 + 1 + 4) + hops * (2 * M + 1) * 4
    peaks = env["peaks"]
    out = {"workload": "hnsw_fp32_l2_1M_d128_M16_efc200_ef64_k10_b256", "graph": "built by the unmodified reference (CPU), bulk-loaded",
           "ids_identical_to_reference": bool(np.array_equal(labels, g["ref_labels"])),
           "scores_identical_to_reference": bool(np.array_equal(scores, g["ref_scores"])),
           "value": nq / (kms * 1e-3), "unit": "queries/s", "ms_per_step": kms,
           "e2e": {"value": nq / float(np.mean(wall)), "unit": "queries/s", "h2d_bytes_per_step": int(Q.nbytes),
                   "d2h_bytes_per_step": int(nq * k * 16)},
           "dist_evals_per_query": evals / nq, "hops_per_query": hops / nq,
           "roofline": {"bound": "hbm", "achieved": touched / (kms * 1e-3) / 1e9, "peak": float(peaks.get("hbm_gbs", 6650.0)),
                        "unit": "GB/s", "frac": touched / (kms * 1e-3) / 1e9 / float(peaks.get("hbm_gbs", 6650.0)), "traffic": None,
                        "kernel": "hnsw_search_kernel", "kernel_ms": kms,
                        "note": "gather-bound (dependent random reads): useful bytes = evals*(row 512 B + link + flag + tag) + hops*link record; no closed form"},
           "graph_load_s": round(load_s, 2), "reference_1core_qps_on_build_host": nq / float(g["ref_query_s_1core"])}
    G.close()
    return out


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
    peaks, peak_src = measured_peaks()
    env = {"torch": torch, "dist": dist, "capi": capi, "sharded": sharded, "rank": rank, "world": world,
           "device": torch.device("cuda", local_rank), "peaks": peaks, "peak_src": peak_src}

    run = FlatRun(env, args.workload, n_total, batch)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = run.time_device(args.steps, args.warmup, args.mode)
    e2e = run.time_e2e(args.steps, args.warmup, args.mode)
    clocks = sampler.stop() if rank == 0 else None
    roof = run.roofline(res, peaks, peak_src) if rank == 0 else None
    frac = run.bounds(res["ms_per_step"], peaks)

    # ---- batch sweep on the headline workload: both roofline fractions at every point ----
    sweep = None
    if args.extras != "none" and world == 1 and not args.batch:
        sweep = []
        for b in (1, 8, 16, 32, 64, 256, 1024):
            if b == run.batch:
                r = res
            else:
                run.set_batch(b)
                r = run.time_device(max(2, min(args.steps, 3)), 2, args.mode)
            pt = {"batch": b, "qps": r["qps"], "ms_per_batch": r["ms_per_step"], "path": run.path_name(r["path"]),
                  "gpu_launches_per_batch": r["launches"] / max(1, (args.steps if b == batch else max(2, min(args.steps, 3))))}
            pt.update({kk: vv for kk, vv in run.bounds(r["ms_per_step"], peaks).items() if kk.endswith("fraction") or kk == "bound"})
            sweep.append(pt)
        run.set_batch(batch)

    cpu_base = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cpu_base = cpu_reference_qps(tname, mname, n_total, dim, k, batch, args.cpu_seconds)
        except Exception as e:  # the baseline is reported, never required
            cpu_base = {"value": None, "unit": "queries/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    n_local, ingest_s, path = run.n_local, run.ingest_s, res["path"]
    path_name = run.path_name(path)
    run.close()

    # ---- the other BASELINE configs as sub-results of the same line ----
    configs = None
    if args.extras != "none" and args.workload == "flat_fp32_ip_10M_d768_k100_b1024" and not (args.rows or args.batch):
        configs = {}
        sub_steps, sub_warm = max(2, min(args.steps, 5)), 3

        def guarded(name, fn):
            try:
                v = fn()
            except Exception as e:  # a sub-result never takes the headline line down
                v = {"error": repr(e)}
            if rank == 0:
                configs[name] = v

        # configs[2]: 6.25 M int8 rows per GPU (the full 50 M x 512 config at 8 GPUs)
        guarded("cfg3_int8_cos_d512_k10_b4096", lambda: sub_config(env, "flat_int8_cos_50M_d512_k10_b4096", 6_250_000 * world,
                                                                    sub_steps, sub_warm, args.mode, peaks))
        if world == 1:
            guarded("cfg4_bf16_ip_20M_d1024_k100_b1024", lambda: sub_config(env, "flat_bf16_ip_20M_d1024_k100_b1024", 20_000_000,
                                                                             sub_steps, sub_warm, args.mode, peaks))
            guarded("cfg5_hnsw_fp32_l2_1M", lambda: cfg5_hnsw(env, max(5, args.steps), 3))
            guarded("cfg1_flat_fp32_l2_100k_single_query", lambda: cfg1_latency(env))

    # ---- N > 1: the sharding inside the C library, one process over all the GPUs (rank 0; the other ranks wait on the CPU) ----
    single = None
    if world > 1 and args.extras != "none" and tname != "fp64":
        cpu_group = dist.new_group(backend="gloo")
        torch.cuda.synchronize()
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                single = single_process_sharded(env, args.workload, n_total, max(3, min(args.steps, 10)), 3, args.mode)
            except Exception as e:
                single = {"error": repr(e)}
        dist.barrier(group=cpu_group)

    if rank == 0:
        line = {
            "metric": METRIC, "value": res["qps"], "unit": "queries/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if tname == "fp32" else tname, "data": "synthetic",
            "config": {"workload": wl_name, "type": tname, "space": mname, "rows": n_total, "rows_per_gpu": n_local, "dim": dim,
                       "k": k, "batch": batch, "path": path_name,
                       "l2_between_iters": "store (%.1f GB/GPU) is larger than L2" % (n_local * dim * ELEM[tname] / 1e9),
                       "sharding": "contiguous row ranges, one all-gather of packed per-shard top-K + merge" if world > 1 else "single GPU",
                       "ingest_s": round(ingest_s, 2)},
            "e2e": e2e, "gpu_launches": int(res["launches"]), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base,
            "hbm_fraction": frac["hbm_fraction"], "tensor_fraction": frac["tensor_fraction"],
            "candidates_per_query": res["candidates"] / max(1, args.steps * batch), "fallback_queries": res["fallbacks"],
            "step_breakdown_ms": {"dominant_kernel": res["scan_ms"], "local_topk_all_kernels": res["local_ms"],
                                  "gather_merge_sync": max(0.0, res["ms_per_step"] - res["local_ms"]), "step": res["ms_per_step"]},
            "wall_s_timed_region": res["wall"], "sweep": sweep, "configs": configs, "single_process": single,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
