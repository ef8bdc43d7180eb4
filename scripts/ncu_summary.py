#!/usr/bin/env python
"""Summarise ncu output into profiles/ (run here, no GPU needed).

    python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_bench.md "<command>"
    python scripts/ncu_summary.py full gpurun_out/prof_apply3d.ncu-rep profiles/r01_apply3d_full.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(src, dst, cmd):
    text = open(src).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    per = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3,
                  "s": 1e6, "second": 1e6}[unit]
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void ", "", name)
        n, t, lo, hi = per.get(name, (0, 0.0, 1e30, 0.0))
        per[name] = (n + 1, t + us, min(lo, us), max(hi, us))
        total += us
    with open(dst, "w") as f:
        f.write(f"# ncu launch list\n\ncommand: `{cmd}`\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv`; per-launch times are "
                "cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                f"{sum(v[0] for v in per.values())} launches, {total / 1e3:.3f} ms of GPU time in total.\n\n"
                "| kernel | launches | total us | avg us | min us | max us | share |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for name, (n, t, lo, hi) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:110]}` | {n} | {t:.1f} | {t / n:.2f} | {lo:.2f} | {hi:.2f} | {100 * t / total:.1f}% |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src.split('/')[-1]}`\n\n"
                "`ncu --set full --clock-control none --import-source on` (one GPU); values per launch.\n\n")
        names = [r[hdr.index("Kernel Name")] for r in data]
        f.write("kernels: " + "; ".join(f"`{n[:90]}`" for n in names) + "\n\n")
        f.write("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |\n")
        f.write("|---|---|" + "---:|" * len(data) + "\n")
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                f.write(f"| `{m}` | {units[i]} | " + " | ".join(r[i] for r in data) + " |\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        full(sys.argv[2], sys.argv[3])
