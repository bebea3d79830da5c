#!/usr/bin/env python3
"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
    python tools/launch_summary.py gpurun_out/launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("tsg::", "")
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print("total %.1f us in %d launches" % (tot, sum(a[0] for a in agg.values())))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-58s %5d %10.1f us %5.1f%%" % (k[:58], c, v, 100 * v / tot))
