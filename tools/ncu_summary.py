#!/usr/bin/env python
"""Summarise ncu output into small text files for profiles/ (run in the build container).

    python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/xxx_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep                 > profiles/xxx_k1_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        a = agg.setdefault(r[kn][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: gpu__time_duration.sum per kernel (ns -> us), {sum(a[0] for a in agg.values())} launches")
    print(f"# {'total us':>10} {'n':>5} {'share':>6} {'avg us':>9}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e3:12.1f} {n:5d} {100 * t / tot:5.1f}% {t / n / 1e3:9.2f}  {k}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"== {r[hdr.index('Kernel Name')]}  (launch id {r[0]})")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:85s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
