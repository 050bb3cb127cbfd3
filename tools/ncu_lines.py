#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export by CUDA source line: stall samples + instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data, fname, hdr = [], "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ws, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
        continue
    if hdr is None or r[0] in ("", "Function Name"):
        continue
    try:
        data.append((int(r[ws]), int(r[ie]), fname, r[0], r[1].strip()[:120]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "total warp-instructions", sum(d[1] for d in data))
for d in sorted(data, reverse=True)[:top]:
    print("%7d %5.1f%% inst=%10d %s:%s  %s" % (d[0], 100.0 * d[0] / tot, d[1], d[2], d[3], d[4]))
