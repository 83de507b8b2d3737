#!/usr/bin/env python
"""Full-recursion timing: index.align() of reveal_b200 (GPU) vs the reference's C aligner (CPU, oracle/_ref),
both driven by the deterministic callbacks of tests/align_callbacks.py (largest MUM first).  Not the bench.py
metric: a secondary measurement of rows a13-a16."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from align_callbacks import make_callbacks  # noqa: E402
from reveal_b200 import reveallib, synth  # noqa: E402
import oracle.ref as R  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2
length = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
minl = int(sys.argv[3]) if len(sys.argv) > 3 else 20
which = sys.argv[4] if len(sys.argv) > 4 else "both"
gs = synth.genomes(ns, length, seed=1)
samples = [[g.tobytes().decode("ascii")] for g in gs]
out = {"genomes": ns, "length": length, "minl": minl}
if which in ("both", "ours"):
    log = []
    idx = reveallib.index()
    for k, s in enumerate(samples):
        idx.addsample("s%d" % k)
        idx.addsequence(s[0])
    t0 = time.perf_counter()
    idx.construct()
    t1 = time.perf_counter()
    mp, ga = make_callbacks(log, minlen=minl)
    idx.align(mp, ga, threads=0, minl=minl, minn=2)
    t2 = time.perf_counter()
    steps = sum(1 for e in log if e[0] == "align")
    aligned = sum(e[1] * e[2] for e in log if e[0] == "align")
    out["ours"] = {"construct_s": t1 - t0, "align_s": t2 - t1, "steps": steps, "aligned_bases": aligned, "aligned_bases_per_s": aligned / (t2 - t0)}
    T_ours = idx.T
if which in ("both", "ref") and R.available():
    log = []
    idx = R.index_from_samples(samples, construct=False)
    t0 = time.perf_counter()
    idx.construct()
    t1 = time.perf_counter()
    mp, ga = make_callbacks(log, minlen=minl)
    idx.align(mp, ga, threads=0, minl=minl, minn=2)
    t2 = time.perf_counter()
    steps = sum(1 for e in log if e[0] == "align")
    aligned = sum(e[1] * e[2] for e in log if e[0] == "align")
    out["reference"] = {"construct_s": t1 - t0, "align_s": t2 - t1, "steps": steps, "aligned_bases": aligned, "aligned_bases_per_s": aligned / (t2 - t0)}
    if which == "both":
        out["identical_text"] = idx.T[:idx.n] == T_ours
print(json.dumps(out))
