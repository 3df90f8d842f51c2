#!/usr/bin/env python
"""Opcode histogram (thread instructions per row/face) + headline metrics of the kernels in an .ncu-rep (run here, no GPU).
    python tools/ncu_ops.py gpurun_out/x.ncu-rep [units_per_launch]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
for r in rows[2:]:
    print({k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "").split(".")[0][-34:]: r[hdr.index(k)][:48]
           for k in keys if k in hdr})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
agg, name, tot = None, None, 0
out = []
for r in csv.reader(src.splitlines()):
    if r and r[0] == "Kernel Name":
        if agg:
            out.append((name, agg, tot))
        agg, name, tot = collections.Counter(), r[1][:70], 0
        continue
    if not r or r[0] == "Address" or len(r) < 7:
        continue
    op = r[1].split()
    m = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
    n = int(r[6])
    agg[m] += n
    tot += n
if agg:
    out.append((name, agg, tot))
for name, agg, tot in out:
    print("==", name, "thread instr", tot, "per unit", round(tot / units, 1))
    print("  ", ", ".join(f"{m} {n / units:.1f}" for m, n in agg.most_common(24)))
