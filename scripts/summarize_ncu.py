#!/usr/bin/env python
"""Turns gpurun_out/*.ncu-rep and launch-list CSVs into the small tracked summaries under profiles/.

  python scripts/summarize_ncu.py full  gpurun_out/prof.ncu-rep  profiles/r1_x_summary.md  "title"
  python scripts/summarize_ncu.py list  gpurun_out/launches.csv  profiles/r1_x_launches.md "title" [skip]
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "gpc__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utcimma_src_int8_sparsity_off.sum.per_second",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "smsp__mem_tensor_reads_op_ldt.sum",
]


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `%s` (ncu --set full --clock-control none --import-source on; numbers taken under the "
                "profiler are for analysis, never bench values).\n\n" % (title, rep))
        for r in rows[2:]:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[idx["Kernel Name"]])
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            f.write("\n")


def launches(path, out, title, skip=0):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("ID")
    data = [(int(r[ii]), r[ki], float(r[vi].replace(",", ""))) for r in rows[hi + 2:] if len(r) > vi]
    data = data[skip:]
    agg = collections.OrderedDict()
    for _, k, v in data:
        name = k.split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in agg.values())
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `%s` (ncu --metrics gpu__time_duration.sum --clock-control none). Per-launch times are "
                "cold-cache and serialised: compare SHARES. First %d launches (ingest / warm-up) skipped.\n\n" % (title, path, skip))
        f.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
        for name, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f%% |\n" % (name, n, v / 1e3, 100 * v / tot))
        f.write("\n## launch list\n\n| id | kernel | us |\n|---|---|---|\n")
        for i, k, v in data:
            f.write("| %d | `%s` | %.1f |\n" % (i, k.split("(")[0], v / 1e3))


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        launches(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]) if len(sys.argv) > 5 else 0)
