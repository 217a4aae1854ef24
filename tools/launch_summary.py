#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total time, share."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
    n = r[ki].split("(")[0]
    agg[n][0] += 1
    agg[n][1] += v * scale
tot = sum(v for _, v in agg.values())
print("%-36s %8s %12s %7s" % ("kernel", "launches", "total_us", "share"))
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-36s %8d %12.1f %6.1f%%" % (n[:36], c, v, 100 * v / tot))
print("%-36s %8d %12.1f" % ("TOTAL", sum(c for c, _ in agg.values()), tot))
