#!/usr/bin/env python
"""One rank's share of a sharded recursion, timed alone in one process (no gather): how long the replicated tree above the cut
takes for world = 1, 2, 4, 8 -- tells an algorithmic cost of the cut apart from contention between the ranks of a box.
usage: shard_prefix_probe.py [genomes] [length]"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reveal_b200 import rem, reveallib, synth  # noqa: E402

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 2
length = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
with tempfile.TemporaryDirectory() as tmp:
    files = []
    for k, g in enumerate(synth.genomes(ng, length, seed=1)):
        files.append(os.path.join(tmp, "g%d.fa" % k))
        with open(files[-1], "w") as f:
            f.write(">g%d\n%s\n" % (k, g.tobytes().decode()))
    for world in (1, 1, 2, 4, 8):
        args = rem.rem_args(files)
        idx = reveallib.index()
        r = rem.Rem(args)
        for fn in files:
            r.read_fasta(fn, idx, contigs=args.contigs, toupper=args.toupper)
        idx.construct()
        with r.recursion_graph():
            mp, ga = r.callbacks(args.minlength)
            t0 = time.perf_counter()
            extra = {"mumpicker_batch": r.batch_picker, "mums_as_rows": True}
            if world > 1:
                extra.update(shard_rank=0, shard_world=world)
            r.shard = None  # (no collection of the other ranks' parts: this probe runs alone)
            idx.align(mp, ga, threads=args.threads, wpen=args.wpen, wscore=args.wscore, minl=args.minlength, minn=args.minn, **extra)
            dt = time.perf_counter() - t0
        st = reveallib.align_stats()
        print(json.dumps({"world": world, "align_s": round(dt, 3), "above_cut_s": round(st["above_cut_s"], 3), "device_step_s": round(st["device_step_s"], 3),
                          "mumpicker_s": round(st["mumpicker_s"], 3), "graphalign_s": round(st["graphalign_s"], 3), "steps": st["steps"], "batches": st["device_batches"],
                          "units": len(idx.shard_units)}))
        del idx
