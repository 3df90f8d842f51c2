#!/usr/bin/env python
"""Selected raw metrics of every kernel in the given .ncu-rep files as one CSV (run here, no GPU):
    python tools/ncu_extract.py out.csv a.ncu-rep b.ncu-rep ..."""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]

out = csv.writer(open(sys.argv[1], "w", newline=""))
out.writerow(["capture"] + KEYS)
units_done = False
for rep in sys.argv[2:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    if not units_done:
        out.writerow(["(unit)"] + [units[hdr.index(k)] if k in hdr else "" for k in KEYS])
        units_done = True
    for r in rows[2:]:
        out.writerow([rep.split("/")[-1]] + [r[hdr.index(k)] if k in hdr else "" for k in KEYS])
