#!/usr/bin/env python
"""Executed warp instructions and stall samples of one kernel of an .ncu-rep (captured with --import-source on) per
source line of raster_render.cu, and the hottest lines.
usage: python scripts/ncu_source_breakdown.py report.ncu-rep render_backward [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname = rows[0][1]
hdr = rows[2]
il, ie, isamp, isrc = hdr.index("Line No"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
per = collections.OrderedDict()
for r in rows[3:]:
    if r and r[0] == "File Name":
        break                                  # the next file's table (inlined headers)
    try:
        per[int(r[il])] = (int(r[ie]), int(r[isamp] or 0), r[isrc])
    except (ValueError, IndexError):
        continue
tot = sum(c for c, _, _ in per.values())
ts = max(1, sum(s for _, s, _ in per.values()))
print(f"{kern}: {fname}: {tot} warp instructions attributed, {ts} stall samples")
print(f"hottest {top_n} source lines (share of executed warp instructions / of stall samples):")
for ln, (c, s, text) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"  line {ln:4d}  {100 * c / tot:5.1f} %  {100 * s / ts:5.1f} %  | {text.strip()[:110]}")
