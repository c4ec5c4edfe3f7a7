#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py launches.csv [steps]"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    name = name.replace("void unnamed>::", "").split("(")[0][:60]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    agg[name][0] += 1
    agg[name][1] += v
    tot += v
print(f"{'us/step':>10} {'launches/step':>14} {'share':>7}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / steps:10.1f} {n / steps:14.1f} {t / tot * 100:6.1f}%  {k}")
print(f"{tot / steps:10.1f} us/step total over {sum(n for n, _ in agg.values()) / steps:.0f} launches/step (cold-cache, serialised under ncu)")
