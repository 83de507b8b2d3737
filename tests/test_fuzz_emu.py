"""Seeded fuzzing of the kernel logic (emulated kernels vs the oracle): random alphabets (incl. IUPAC, lower case,
one-letter), 1-5 samples, empty samples and contigs, periodic / mutated / unrelated sequences, rc, random minl / minn;
checks SA, SAi, LCP, SO and all four sweeps (getmums, getmums_rem, getmultimums, getmultimems) bit-exactly.
A longer unbounded version of the same loop found the rc-with-empty-first-sample underrun (now refused)."""
import ctypes

import numpy as np
import pytest

import oracle.port as P
from reveal_b200 import _native
from util import NativeIndex, assert_same

ALPHABETS = [b"A", b"AC", b"ACGT", b"ACGTN", b"ACGTNacgt", b"ACGTRYKMSWBDHVN", b"ACGTRYKMSWBDHVNacgtn", b"AN"]


def random_case(rng):
    al = np.frombuffer(ALPHABETS[rng.integers(len(ALPHABETS))], np.uint8)
    ns = int(rng.integers(1, 6))
    mode = int(rng.integers(4))
    base = al[rng.integers(0, len(al), size=int(rng.integers(1, 300)))]
    samples = []
    for _ in range(ns):
        ncont = int(rng.integers(0, 4)) if ns > 1 else int(rng.integers(1, 4))
        seqs = []
        for _ in range(ncont):
            if mode == 0:
                q = al[rng.integers(0, len(al), size=int(rng.integers(0, 200)))]
            elif mode == 1:
                q = base.copy()
                m = rng.random(len(q)) < 0.05
                q[m] = al[rng.integers(0, len(al), size=int(m.sum()))]
            elif mode == 2:
                q = np.tile(al[rng.integers(0, len(al), size=int(rng.integers(1, 6)))], int(rng.integers(1, 80)))
            else:
                a = int(rng.integers(0, len(base)))
                q = base[a:int(rng.integers(a, len(base) + 1))]
            seqs.append(q.tobytes())
        samples.append(seqs)
    return samples


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_emulated_kernels_vs_oracle(emu_lib, seed):
    rng = np.random.default_rng(1000 + seed)
    done = 0
    while done < 6:
        samples = random_case(rng)
        T, nsep, _ = P.assemble(samples)
        if len(T) == 0:
            continue
        ns = len(samples)
        minl, minn = int(rng.integers(0, 12)), int(rng.integers(2, 4))
        rc = int(rng.integers(2)) if (ns >= 2 and nsep[0] >= 0) else 0
        o = P.Index(T, nsep, ns, rc)
        with NativeIndex(emu_lib, T, nsep, ns, rc=rc) as idx:
            assert_same(idx.arr("SA"), o.SA, "SA")
            assert_same(idx.arr("SAi"), o.SAi, "SAi")
            assert_same(idx.arr("LCP"), o.LCP, "LCP")
            if ns > 2:
                assert_same(idx.arr("SO"), o.SO, "SO")
            if ns >= 2:
                for fl in (0, 1):
                    assert_same(idx.mums(minl, fl), o.getmums(minl, rem=bool(fl)), "getmums")
            h, m = idx.multimums(minl, minn)
            oh, om = o.getmultimums(minl, minn)
            assert_same(h, oh, "getmultimums hdr")
            assert_same(m, om, "getmultimums members")
            if ns >= 2:
                nr, nm = ctypes.c_int64(), ctypes.c_int64()
                _native.check(emu_lib, emu_lib.rv_mems_multi_count(idx.h, minl, minn, ctypes.byref(nr), ctypes.byref(nm)))
                hh, mm = np.empty((nr.value, 3), np.int64), np.empty((nm.value, 2), np.int64)
                _native.check(emu_lib, emu_lib.rv_mums_multi_fetch(idx.h, hh.ctypes.data, nr.value, mm.ctypes.data, nm.value))
                oh, om = o.getmultimems(minl, minn)
                assert_same(hh, oh, "getmultimems hdr")
                assert_same(mm, om, "getmultimems members")
        done += 1


def test_rc_with_empty_first_sample_is_refused(emu_lib):
    T, nsep, _ = P.assemble([[], ["ACGT"]])
    with pytest.raises(_native.NativeError):
        NativeIndex(emu_lib, T, nsep, 2, rc=1)


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_emulated_dna_with_rare_symbols(emu_lib, seed):
    """The `mid` leg of scripts/gpu_fuzz.py scaled down to emulator sizes: related DNA genomes with repeats, tandem arrays, N runs,
    stray IUPAC letters and many contigs -- the texts whose keys are 2-bit digits with rare symbols ending them (KeyBlock4 and
    KeyRoller side by side: windows inside the text take the first, the last 32 positions the second)."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("gpu_fuzz", os.path.join(os.path.dirname(here), "scripts", "gpu_fuzz.py"))
    Z = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(Z)
    Z.SCALE = 0.02
    for trial in range(4):
        Z.leg_mid(emu_lib, np.random.default_rng(4200 + 10 * seed + trial))
