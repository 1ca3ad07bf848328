#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics from the raw page and, from the
SASS source page, executed warp instructions by opcode and the warp-stall sample distribution.
usage: ncu_summary.py REPORT.ncu-rep [units_per_launch]   (units = e.g. stored modes -> per-32-modes figures)"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, unit, vals)}
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__cycles_elapsed.avg.per_second", "lts__t_sector_hit_rate.pct"]
for k in keys:
    if k in m:
        print(f"{k:70s} {m[k][0]} {m[k][1]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ci = {name: i for i, name in enumerate(h)}
ops, stalls = Counter(), Counter()
total = 0
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
for r in rows[2:]:
    if len(r) < len(h):
        continue
    sass = r[ci["Source"]].strip()
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    n = int(r[ci["Instructions Executed"]] or 0)
    ops[op.split(".")[0]] += n
    total += n
    for s in stall_cols:
        stalls[s] += int(r[ci[s]] or 0)
per = (units / 32.0) if units else None
print(f"\nexecuted warp instructions: {total}" + (f"  = {total / per:.1f} per 32 units" if per else ""))
for op, n in ops.most_common(28):
    print(f"  {op:14s} {n:>14d} {100.0 * n / total:6.2f}%" + (f"  {n / per:7.2f}/32u" if per else ""))
tot_s = sum(stalls.values()) or 1
print("\nwarp stall samples:")
for s, n in stalls.most_common(10):
    print(f"  {s:28s} {100.0 * n / tot_s:6.2f}%")
