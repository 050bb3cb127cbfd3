#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and every conv launch."""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
tot, agg = 0.0, {}
for r in rows:
    n = r["Kernel Name"].split("(")[0][-44:]
    t = float(r["Metric Value"]) / 1e3
    tot += t
    a = agg.setdefault(n, [0.0, 0])
    a[0] += t
    a[1] += 1
print("total us %.1f over %d launches" % (tot, len(rows)))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-46s %8.1f us %5.1f%% %3d launches" % (k, v[0], 100 * v[0] / tot, v[1]))
if len(sys.argv) > 2:
    i = 0
    for r in rows:
        if sys.argv[2] in r["Kernel Name"]:
            print(i, r["Kernel Name"].split("(")[0][-28:], r["Grid Size"], "%.1f us" % (float(r["Metric Value"]) / 1e3))
            i += 1
