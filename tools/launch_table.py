#!/usr/bin/env python
"""Summarise an ncu launch list with several metrics per launch (gpu__time_duration.sum, dram__bytes_read/write.sum,
sm__pipe_tensor_cycles_active...): per-kernel-function totals and one line per launch of the LAST forward in the list.
usage: python tools/launch_table.py <launches.csv> [launches per forward]"""
import csv
import re
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = int(sys.argv[2]) if len(sys.argv) > 2 else 44
launch = OrderedDict()
for r in csv.DictReader(lines):
    d = launch.setdefault(int(r["ID"]), {"name": r["Kernel Name"], "grid": r["Grid Size"]})
    d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))


def short(n):
    n = re.sub(r"\(sddm::.*", "", n)
    n = n.replace("void sddm::<unnamed>::", "").replace("sddm::<unnamed>::", "").replace("sddm::", "")
    return n[:60]


ids = list(launch)
last = ids[-per:]
tot = sum(launch[i]["gpu__time_duration.sum"] for i in last) / 1e3
print("last forward of the list: %d launches, %.1f us (ncu: cold cache, serialised launches)" % (len(last), tot))
agg = OrderedDict()
for i in last:
    d = launch[i]
    k = re.sub(r"<.*", "", short(d["name"]))
    a = agg.setdefault(k, [0.0, 0, 0.0])
    a[0] += d["gpu__time_duration.sum"] / 1e3
    a[1] += 1
    a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-28s %8.1f us %5.1f%% %3d launches  %7.1f MB dram" % (k, v[0], 100 * v[0] / tot, v[1], v[2] / 1e6))
print("%3s %-62s %-12s %9s %9s %9s %7s" % ("#", "kernel", "grid", "us", "dram MB", "GB/s", "tensor%"))
for n, i in enumerate(last):
    d = launch[i]
    us = d["gpu__time_duration.sum"] / 1e3
    mb = (d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6
    print("%3d %-62s %-12s %9.1f %9.1f %9.0f %7.1f" % (n, short(d["name"]), d["grid"], us, mb, mb / us * 1e3 if us else 0,
                                                     d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)))
