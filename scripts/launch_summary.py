#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    if u in ("nsecond", "ns"):
        v /= 1e3
    elif u in ("msecond", "ms"):
        v *= 1e3
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| %s | %d | %.1f | %.1f | %.1f %% |" % (k[:70], c, t, t / c, 100 * t / tot))
print("total us: %.1f" % tot)
