#!/usr/bin/env python
"""Headline counters of every launch in an .ncu-rep (read here, no GPU needed) and, with --launches, the per-kernel
shares of an ncu launch-list csv.
usage: python scripts/ncu_step_kernels.py report.ncu-rep ["title"]      -> text on stdout (profiles/*_step_kernels.txt)
       python scripts/ncu_step_kernels.py --launches launches.csv       -> markdown table rows on stdout"""
import collections
import csv
import io
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")


def short(name: str) -> str:
    return name.replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")


def report(rep: str, title: str) -> None:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ik = hdr.index("Kernel Name")
    print(f"==================== {title} (ncu --set full --clock-control none; one launch of each kernel)")
    for r in rows[2:]:
        print(short(r[ik]))
        for h, u, v in zip(hdr, units, r):
            if h in KEEP or ("issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v.replace(",", "") or 0) > 0.3):
                print(f"    {h} [{u}] = {v}")


def launches(path: str) -> None:
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
    t, n = collections.OrderedDict(), collections.Counter()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"]).split("(")[0][:70]
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else v
        t[k] = t.get(k, 0.0) + v
        n[k] += 1
    ours = ("preprocess", "scan_tiles", "scatter", "sort_pack", "render_")
    step = sum(v / n[k] for k, v in t.items() if k.startswith(ours))
    print("| kernel | launches | mean us / launch | share of the rasterizer step |\n|---|---|---|---|")
    for k, v in t.items():
        share = f"{100 * v / n[k] / step:.1f} %" if k.startswith(ours) else "(set-up / torch)"
        print(f"| `{k}` | {n[k]} | {v / n[k]:.1f} | {share} |")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
