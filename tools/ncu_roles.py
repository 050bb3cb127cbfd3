#!/usr/bin/env python
"""Per-role busy fraction of conv_tc.cu from an ncu cuda,sass source export: samples inside a role's line range that are
NOT in mbarrier waits, per warp of the role, relative to the samples one always-resident warp receives."""
import csv, re, sys
src = open(sys.argv[2]).read().split("\n")
marks = {}
for i, l in enumerate(src, 1):
    for key, pat in (("epilogue", "====== epilogue"), ("mma", "====== MMA issuer"), ("wload", "====== weight loader"),
                     ("tma", "====== raw-slab TMA issuer"), ("xform", "====== transform"), ("end", "tc_fence_before();\n")):
        if pat.strip() in l and key not in marks:
            marks[key] = i
# end of roles: the teardown
for i, l in enumerate(src, 1):
    if "__syncthreads();" in l and i > marks["xform"]:
        marks["end"] = i - 1
        break
order = sorted(marks.items(), key=lambda kv: kv[1])
rows = list(csv.reader(open(sys.argv[1])))
fname, hdr, data = "", None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; ws = hdr.index("Warp Stall Sampling (All Samples)"); continue
    if hdr is None or r[0] in ("", "Function Name"): continue
    try: data.append((fname, int(r[0]), int(r[ws]), r[1]))
    except (ValueError, IndexError): pass
tot = sum(d[2] for d in data)
waits = sum(d[2] for d in data if d[0] == "conv_tc.cu" and ("mbar_try" in d[3] or "spins" in d[3]))
print("total samples", tot, "in mbarrier waits", waits)
nwarps = {"epilogue": 4, "mma": 1, "wload": 1, "tma": 1, "xform": 8}
per_warp = tot / 16.0
for (k, lo), (_, hi) in zip(order[:-1], order[1:]):
    s = sum(d[2] for d in data if d[0] == "conv_tc.cu" and lo <= d[1] < hi)
    print("%-9s lines %d-%d: %6d samples in role body (excludes inlined helpers), %.2f of one warp's timeline per warp" % (k, lo, hi, s, s / nwarps[k] / per_warp))
