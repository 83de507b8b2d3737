#!/usr/bin/env python
"""Where a step through the drop-in spends its time: index() / addsample+addsequence / construct() / getmums() (C2)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from reveal_b200 import reveallib  # noqa: E402

T, nsep, ns, _ = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c2", 1)
n = len(T)
bounds = [0] + [int(x) + 1 for x in nsep] + [n]
seqs = [T[bounds[k]:bounds[k + 1] - 1].tobytes().decode("ascii") for k in range(ns)]
acc = {"index": 0.0, "addsequence": 0.0, "construct": 0.0, "getmums": 0.0, "dealloc": 0.0}
reps = 8
for rep in range(reps + 2):
    t0 = time.perf_counter()
    idx = reveallib.index()
    t1 = time.perf_counter()
    for k, s in enumerate(seqs):
        idx.addsample("g%d" % k)
        idx.addsequence(s)
    t2 = time.perf_counter()
    idx.construct()
    t3 = time.perf_counter()
    mums = idx.getmums(20) if ns == 2 else idx.getmultimums(minlength=20, minn=2)
    t4 = time.perf_counter()
    del idx, mums
    t5 = time.perf_counter()
    if rep >= 2:
        for key, dt in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
            acc[key] += dt
print(json.dumps({k: round(1e3 * v / reps, 3) for k, v in acc.items()}))
