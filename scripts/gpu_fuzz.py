#!/usr/bin/env python
"""Time-bounded seeded fuzzing of the CUDA path on a real GPU (test infrastructure; run under gpurun).

Four legs, every one compares bit-exactly and stops at the first difference (seed and parameters are printed):
  tiny    the random cases of tests/test_fuzz_emu.py (odd alphabets, empty samples / contigs, periodic text, rc):
          SA, SAi, LCP, SO and all four sweeps against the oracle port
  mid     2-5 related genomes of 3 kbp - 400 kbp with repeats, tandem arrays, N runs, IUPAC letters, many contigs:
          arrays + sweeps against the oracle port (these sizes walk the group / cap / stage-4 thresholds)
  align   index.align() against the reference's unmodified C aligner (oracle/_ref) under the same deterministic callbacks,
          in the reference's order (threads=1) and with frontier batching (threads=0)
  rem     FASTA files -> alignment graph through reveal_b200/rem.py, this library against the reference's extension under the
          same driver, random option mixes: canonical graphs must be identical

usage: gpu_fuzz.py [seconds per leg = 30] [seed = 1] [legs = tiny,mid,align,rem]
Prints one JSON line; exit code 1 on a mismatch.  RV_FUZZ_SCALE (default 1.0) scales the input sizes."""
import json
import os
import sys
import tempfile
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402

import oracle.port as P  # noqa: E402
import oracle.ref as R  # noqa: E402
from reveal_b200 import _native, rem, synth  # noqa: E402
import test_fuzz_emu as F  # noqa: E402
import test_align as A  # noqa: E402
import make_rem_golden as M  # noqa: E402
from util import check_against_oracle, random_related  # noqa: E402

IUPAC = np.frombuffer(b"RYKMSWBDHVN", np.uint8)
SCALE = float(os.environ.get("RV_FUZZ_SCALE", "1"))


def leg_tiny(L, rng):
    samples = F.random_case(rng)
    T, nsep, _ = P.assemble(samples)
    if len(T) == 0:
        return None
    ns = len(samples)
    minl, minn = int(rng.integers(0, 12)), int(rng.integers(2, 4))
    rc = int(rng.integers(2)) if (ns >= 2 and nsep[0] >= 0) else 0
    check_against_oracle(L, T, nsep, ns, rc=rc, minl=minl, minn=minn)
    return {"n": int(len(T)), "ns": ns, "rc": rc}


def mid_samples(rng):
    ns = int(rng.integers(2, 6))
    length = max(200, int(SCALE * 10 ** rng.uniform(3.5, 5.6)))
    seed = int(rng.integers(1, 1 << 30))
    snp, indel = float(rng.choice([0.0, 0.002, 0.01, 0.05])), float(rng.choice([0.0, 0.001, 0.004]))
    kind = int(rng.integers(4))
    if kind == 0:
        gs = synth.genomes(ns, length, seed=seed, snp=snp, indel=indel)
    elif kind == 1:
        gs = synth.repeat_genomes(ns, length, seed=seed, snp=snp, indel=indel, n_runs=int(rng.integers(0, 6)))
    elif kind == 2:   # low-complexity: a short unit repeated with rare changes (long matches, deep groups, stage 4)
        unit = rng.integers(0, 4, size=int(rng.integers(1, 400)), dtype=np.uint8)
        base = np.tile(unit, length // len(unit) + 1)[:length]
        gs = [synth._ACGT[base if k == 0 else synth.mutate(base, np.random.default_rng(seed + k), snp / 10, indel / 10)] for k in range(ns)]
    else:             # identical copies and copies of pieces
        g0 = synth.genomes(1, length, seed=seed)[0]
        gs = [g0]
        for k in range(1, ns):
            a = int(rng.integers(0, length // 2))
            gs.append(g0[a:int(rng.integers(a + 1, length + 1))].copy())
    gs = [np.array(g, dtype=np.uint8) for g in gs]
    if rng.random() < 0.4:   # stray IUPAC letters / lower case
        for g in gs:
            k = int(rng.integers(0, 1 + len(g) // 500))
            g[rng.integers(0, len(g), size=k)] = IUPAC[rng.integers(0, len(IUPAC), size=k)]
    samples = []
    for g in gs:
        if rng.random() < 0.5:
            cuts = np.unique(rng.integers(1, max(2, len(g)), size=int(rng.integers(1, 2 + min(3000, len(g) // 20)))))
            samples.append([p for p in np.split(g, cuts)])
        else:
            samples.append([g])
    return samples


def leg_mid(L, rng):
    samples = mid_samples(rng)
    T, nsep = synth.concat(samples)
    ns = len(samples)
    minl, minn = int(rng.integers(5, 40)), int(rng.integers(2, ns + 1))
    rc = int(rng.integers(2)) if rng.random() < 0.3 else 0
    check_against_oracle(L, T, nsep, ns, rc=rc, minl=minl, minn=minn)
    return {"n": int(len(T)), "ns": ns, "rc": rc}


def leg_align(reveallib, rng):
    ns = int(rng.integers(2, 6))
    length = max(200, int(SCALE * 10 ** rng.uniform(3.0, 4.9)))
    sigma = int(rng.choice([2, 3, 4, 4, 4]))
    minl, minn = int(rng.integers(6, 24)), int(rng.integers(2, ns + 1))
    if rng.random() < 0.35:   # repeat in front of a match: the bubble_sort replay paths
        samples = A.repeat_before_match(rng, rlen=int(rng.integers(50, 3000)), mlen=int(rng.integers(100, 2000)), nsamples=min(ns, 3))
    else:
        samples = random_related(rng, ns, length, sigma=sigma, snp=float(rng.choice([0.005, 0.02, 0.08])))
    maxsteps = None if rng.random() < 0.7 else int(rng.integers(1, 200))
    ref = A.run_reference(samples, minl, minn, maxsteps)
    steps = A.compare(ref, A.run_ours(reveallib, samples, minl, minn, maxsteps, threads=1), ordered=True)
    if maxsteps is None:   # a step budget makes the visited set depend on the visiting order
        A.compare(ref, A.run_ours(reveallib, samples, minl, minn, None, threads=0), ordered=False)
    return {"ns": len(samples), "steps": steps}


def leg_rem(reveallib, rng, tmp, trial):
    ng = int(rng.integers(2, 6))
    length = max(300, int(SCALE * 10 ** rng.uniform(3.2, 5.3)))
    seed = int(rng.integers(1, 1 << 30))
    if rng.random() < 0.3:
        gs = synth.repeat_genomes(ng, max(length, int(20000 * SCALE)), seed=seed, snp=0.01, indel=0.002, n_runs=int(rng.integers(0, 3)))
    else:
        gs = synth.genomes(ng, length, seed=seed, snp=float(rng.choice([0.002, 0.01, 0.04])), indel=float(rng.choice([0.0, 0.002])))
    files = []
    for k, g in enumerate(gs):
        files.append(os.path.join(tmp, "f%d_g%d.fa" % (trial, k)))
        M.write_fasta(files[-1], "g%d" % k, np.asarray(g, np.uint8).tobytes().decode())
    opts = dict(minlength=int(rng.integers(8, 24)), minn=int(rng.integers(2, ng + 1)), trim=bool(rng.integers(0, 2)),
                seedsize=int(rng.choice([0, 15, 10000])), maxmums=int(rng.choice([2, 5, 1000, 1000])), wpen=int(rng.integers(0, 4)),
                wscore=int(rng.integers(1, 4)), gcmodel=str(rng.choice(["sumofpairs", "star-avg", "star-med"])))
    got = []
    for module in (R.module(32), reveallib):
        G, idx = rem.align_genomes(rem.rem_args(files, **opts), index_module=module)
        T = idx.T
        if ng > 2:
            rem.prune_nodes(G, T=T)
        got.append(M.canonical(G, T))
    for f in files:
        os.unlink(f)
    if got[0] != got[1]:
        raise AssertionError("alignment graphs differ: %r" % (opts,))
    return {"ng": ng, "nodes": len(got[0]["nodes"])}


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    legs = (sys.argv[3] if len(sys.argv) > 3 else "tiny,mid,align,rem").split(",")
    from reveal_b200 import reveallib
    emu = os.environ.get("RV_FUZZ_EMU_LIB")   # dry run of this script in the GPU-less container: the emulated kernels (tests/emu)
    if emu:
        L = _native.bind(emu)
        reveallib._load(emu)                  # needs REVEAL_B200_TEST_HOOKS=1
    else:
        L = _native.lib()
    out = {"seed": seed0, "seconds_per_leg": budget, "legs": {}}
    failed = False
    with tempfile.TemporaryDirectory() as tmp:
        for leg in legs:
            rng_seed = seed0 * 1000 + {"tiny": 1, "mid": 2, "align": 3, "rem": 4}[leg]
            t0 = time.perf_counter()
            cases, biggest, trial = 0, 0, 0
            rec = {"cases": 0, "failed": None}
            while time.perf_counter() - t0 < budget:
                case_seed = rng_seed * 100000 + trial
                rng = np.random.default_rng(case_seed)
                trial += 1
                try:
                    if leg == "tiny":
                        r = leg_tiny(L, rng)
                    elif leg == "mid":
                        r = leg_mid(L, rng)
                    elif leg == "align":
                        r = leg_align(reveallib, rng)
                    else:
                        r = leg_rem(reveallib, rng, tmp, trial)
                except Exception as e:  # a mismatch (AssertionError) or a native error: both are findings
                    rec["failed"] = {"case_seed": case_seed, "error": "%s: %s" % (type(e).__name__, str(e)[:600]), "trace": traceback.format_exc()[-1200:]}
                    failed = True
                    break
                if r is not None:
                    cases += 1
                    biggest = max(biggest, r.get("n", r.get("steps", r.get("nodes", 0))))
            rec["cases"] = cases
            rec["largest"] = biggest
            rec["seconds"] = round(time.perf_counter() - t0, 1)
            out["legs"][leg] = rec
    out["ok"] = not failed
    print(json.dumps(out))
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
