#!/usr/bin/env python
"""Reads an .ncu-rep (here, no GPU needed): SASS opcode histogram weighted by executed count + the headline counters.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]
ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot, samp, n = collections.Counter(), collections.Counter(), 0
for r in rows[hi + 1:]:
    try:
        c = int(r[ie])
    except Exception:
        continue
    toks = r[ia].strip().split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    tot[op] += c
    n += c
    try:
        samp[op] += int(r[isamp])
    except Exception:
        pass
print(rows[0][1][:120] if rows and len(rows[0]) > 1 else "")
print("warp instructions executed:", n)
for op, c in tot.most_common(24):
    print(f"  {op:10s} {c:12d} {100 * c / n:5.1f}%  stall samples {samp[op]}")
rr = list(csv.reader(io.StringIO(raw)))
keep = ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
for h, u, v in zip(rr[0], rr[1], rr[-1]):
    if h in keep or ("stalled" in h and "per_issue_active" in h and float(v or 0) > 0.05):
        print(f"  {h} [{u}] = {v}")
