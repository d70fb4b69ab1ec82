#!/usr/bin/env python
"""Per-kernel average duration and share from an ncu launch list
(ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...).  Usage: launch_summary.py X.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, agg = None, collections.defaultdict(list)
for r in rows:
    if len(r) > 5 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        if d["Metric Name"] == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            if d["Metric Unit"] == "ns":
                v /= 1000
            agg[d["Kernel Name"][:70]].append(v)
tot = sum(sum(v) for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"{k:72s} n={len(v):3d} avg={sum(v) / len(v):8.2f}us share={100 * sum(v) / tot:5.1f}%")
