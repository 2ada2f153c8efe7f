#!/usr/bin/env python3
"""Instructions executed and stall samples per CUDA source line of a kernel, straight from an .ncu-rep captured
with --import-source on (no nvdisasm needed, works for a build that no longer exists):
    python scripts/ncu_cuda_lines.py <report.ncu-rep> [top]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg = defaultdict(lambda: [0, 0, ""])
fname, hdr = "", None
for row in csv.reader(io.StringIO(out)):
    if len(row) == 2 and row[0] == "File Path":
        fname = row[1].split("/")[-1]
        continue
    if row and row[0] == "Line No":
        hdr = row
        ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(row) <= max(ie, si):
        continue
    line = row[0]
    try:
        n, s = int(row[ie] or 0), int(row[si] or 0)
    except ValueError:
        continue
    if row[1].strip():
        agg[(fname, line)][2] = row[1].strip()[:110]
    agg[(fname, line)][0] += n
    agg[(fname, line)][1] += s
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("total warp instructions", tot, "samples", ts)
for (f, l), (n, s, txt) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n / tot * 100:5.1f}% inst {s / ts * 100:5.1f}% samp  {f}:{l}  {txt}")
