"""Kernel LOGIC check without a GPU: the product's .cu sources compiled for the
host with the fiber emulation of tests/emu (same C-ABI), compared bit-exactly
with the golden vectors and the oracle.  The real parity gate is the gpu-marked
suite; this tier exists so that indexing/scan/look-back bugs are caught here."""
import numpy as np
import pytest

import oracle.port as P
from conftest import golden_names, load_golden
from util import check_handle_reuse, check_pack_block, check_sweep_prefetch, check_against_golden, check_against_oracle, random_related


@pytest.mark.parametrize("name", golden_names())
def test_emu_matches_golden(emu_lib, name):
    check_against_golden(emu_lib, load_golden(name))


@pytest.mark.parametrize("nsamples,length,sigma,minl", [(2, 1500, 2, 4), (3, 3000, 4, 6), (2, 9000, 4, 8), (5, 2500, 3, 5)])
def test_emu_matches_oracle_random(emu_lib, nsamples, length, sigma, minl):
    rng = np.random.default_rng(nsamples * 1000 + length)
    T, nsep, _ = P.assemble(random_related(rng, nsamples, length, sigma))
    check_against_oracle(emu_lib, T, nsep, nsamples, minl=minl)


def test_emu_single_sample_and_tiny(emu_lib):
    for text in (b"A", b"AC", b"ACA", b"GATTACA"):
        T, nsep, _ = P.assemble([[text]])
        check_against_oracle(emu_lib, T, nsep, 1, minl=0)


def test_emu_rc(emu_lib):
    rng = np.random.default_rng(7)
    s = random_related(rng, 2, 2000, 4)
    T, nsep, _ = P.assemble(s)
    check_against_oracle(emu_lib, T, nsep, 2, rc=1, minl=6)


def test_emu_long_identical_stretch_defers_to_doubling(emu_lib):
    """Matches longer than the comparison cap (SA_CMP_CAP) must fall back to the doubling rounds."""
    rng = np.random.default_rng(3)
    al = np.frombuffer(b"ACGT", np.uint8)
    s = al[rng.integers(0, 4, size=7000)].tobytes()
    t = bytearray(s)
    t[6500] = ord("A") if t[6500] != ord("A") else ord("C")
    T, nsep, _ = P.assemble([[s], [bytes(t)], [s[100:6900]]])
    check_against_oracle(emu_lib, T, nsep, 3, minl=10)


def test_emu_large_groups_and_small_groups_mixed(emu_lib):
    rng = np.random.default_rng(4)
    al = np.frombuffer(b"ACGT", np.uint8)
    rep = al[rng.integers(0, 4, size=60)].tobytes()
    s0 = al[rng.integers(0, 4, size=1500)].tobytes() + rep * 25 + al[rng.integers(0, 4, size=1500)].tobytes()
    s1 = s0[:700] + b"G" + s0[701:2000] + rep * 3 + s0[2000:]
    T, nsep, _ = P.assemble([[s0], [s1]])
    check_against_oracle(emu_lib, T, nsep, 2, minl=8)


@pytest.mark.parametrize("bits", ["32", "64"])
def test_emu_both_key_widths(emu_lib, monkeypatch, bits):
    monkeypatch.setenv("RV_SA_KEY_BITS", bits)
    rng = np.random.default_rng(21)
    T, nsep, _ = P.assemble(random_related(rng, 3, 2500, 4))
    check_against_oracle(emu_lib, T, nsep, 3, minl=6)
    check_against_golden(emu_lib, load_golden("tandem"))
    check_against_golden(emu_lib, load_golden("all_A"))


def test_emu_byte_comparison_path(emu_lib, monkeypatch):
    """RV_SA_NO_PACK forces the byte-wise comparison loops that large alphabets use."""
    monkeypatch.setenv("RV_SA_NO_PACK", "1")
    rng = np.random.default_rng(33)
    T, nsep, _ = P.assemble(random_related(rng, 3, 2500, 4))
    check_against_oracle(emu_lib, T, nsep, 3, minl=6)
    check_against_golden(emu_lib, load_golden("with_N_d2"))


def test_emu_handle_reuse_alphabet_cache(emu_lib):
    check_handle_reuse(emu_lib)


def test_emu_sweeps_between_rebuilds(emu_lib):
    check_sweep_prefetch(emu_lib)


def test_emu_pack_into_peer_block(emu_lib):
    check_pack_block(emu_lib)


@pytest.mark.parametrize("ngenomes,length", [(2, 40000), (5, 15000)])
def test_emu_dna_at_three_digit_passes(emu_lib, ngenomes, length):
    """Synthetic genomes large enough for 12-mer keys (n > 65 536): three byte-wide digit passes, the DNA key block (two 128-bit text
    loads, 2-bit packed window) in the histogram kernel and the first pass, the byte-digit histogram -- the configuration of
    BASELINE configs[1] -- with N runs and stray IUPAC letters as rare symbols; arrays and sweeps against the oracle."""
    from reveal_b200 import synth
    rng = np.random.default_rng(17 + ngenomes)
    gs = [np.array(g) for g in synth.genomes(ngenomes, length, seed=31 + ngenomes, snp=0.01, indel=0.001)]
    for g in gs:
        at = int(rng.integers(100, length - 200))
        g[at:at + int(rng.integers(1, 40))] = ord("N")
        g[rng.integers(0, len(g), size=3)] = np.frombuffer(b"RYK", np.uint8)
    T, nsep = synth.concat([[g] for g in gs])
    check_against_oracle(emu_lib, T, nsep, ngenomes, minl=12, minn=2)
