#!/usr/bin/env python
"""End-to-end `rem`: FASTA files -> alignment graph (construct + align + prune), aligned bases per second as the
reference reports them (reveal/rem.py:470-490), SURVEY.md 8(d)(ii).

Two arms with the SAME driver (reveal_b200/rem.py): the B200 library behind `reveal_b200.reveallib`, and the
reference's own compiled extension (oracle/_ref, CPU) -- so the ratio isolates what the index path buys end to end.
Not the bench.py metric: a secondary measurement.

usage: rem_bench.py c1|c2|<n_genomes> [length] [ours|ref|both]"""
import gzip
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reveal_b200 import rem, synth  # noqa: E402


def write_fasta(path, name, seq):
    with open(path, "w") as f:
        f.write(">%s\n" % name)
        for i in range(0, len(seq), 100):
            f.write(seq[i:i + 100] + "\n")


def run(files, module, label):
    args = rem.rem_args(files)
    t0 = time.perf_counter()
    G, idx = rem.align_genomes(args, index_module=module)
    t1 = time.perf_counter()
    T = idx.T
    if len(G.graph["paths"]) > 2:
        rem.prune_nodes(G, T=T)
    t2 = time.perf_counter()
    bases, total, nodes = rem.aligned_bases(G, idx)
    stats = module.align_stats() if hasattr(module, "align_stats") else None
    try:
        from reveal_b200 import remcore
        chain = remcore.chain_stats()
        chain["pick_phases_s"] = remcore.pick_phases()
    except Exception:
        chain = None
    return {"arm": label, "align_stats": stats, "chain_stats": chain, "seconds": t2 - t0, "align_genomes_s": t1 - t0, "aligned_bases": bases, "total_bases": total,
            "aligned_nodes": nodes, "nodes": G.number_of_nodes(), "aligned_bases_per_s": bases / (t2 - t0)}, T


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "c1"
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    with tempfile.TemporaryDirectory() as tmp:
        files = []
        if what == "c1":  # BASELINE configs[0]: the reference's tests/1a.fa + 1b.fa (sequences kept with the golden graphs)
            inputs = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "rem", "inputs.json.gz")).read())
            for fn in ("1a.fa", "1b.fa"):
                files.append(os.path.join(tmp, fn))
                with open(files[-1], "w") as f:
                    for name, seq in inputs[fn]:
                        f.write(">%s\n%s\n" % (name, seq))
            desc = "1a.fa + 1b.fa (BASELINE configs[0])"
        else:
            ng = 2 if what == "c2" else int(what)
            length = int(sys.argv[2]) if len(sys.argv) > 2 and what != "c2" else 5_000_000
            for k, g in enumerate(synth.genomes(ng, length, seed=1)):
                files.append(os.path.join(tmp, "g%d.fa" % k))
                write_fasta(files[-1], "g%d" % k, g.tobytes().decode())
            desc = "%d synthetic genomes of %d bp (1%% SNP, 0.1%% indel)" % (ng, length)
        out = {"workload": desc, "options": "rem defaults (-m 20 -n 2)"}
        texts = {}
        if which in ("both", "ours"):
            from reveal_b200 import reveallib
            run(files, reveallib, "warm-up")  # CUDA context, module load
            out["ours"], texts["ours"] = run(files, reveallib, "reveal_b200.reveallib (B200)")
        if which in ("both", "ref"):
            import oracle.ref as R
            if R.available():
                out["reference"], texts["ref"] = run(files, R.module(32), "reference extension (CPU, oracle/_ref), same driver")
        if len(texts) == 2:
            out["identical_text"] = texts["ours"] == texts["ref"]
            out["identical_aligned_bases"] = out["ours"]["aligned_bases"] == out["reference"]["aligned_bases"]
            out["speedup"] = out["reference"]["seconds"] / out["ours"]["seconds"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
