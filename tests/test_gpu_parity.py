"""Parity tests proper: the sm_100a CUDA path, called through the C-ABI of
libreveal_b200.so on a real GPU, against (1) the golden vectors minted from the
unmodified reference, (2) the CPU oracle on the same seeded inputs, and (3) at
BASELINE sizes through size-independent properties.  Everything is integer/byte
work: the bar is bit-exact equality."""
import numpy as np
import pytest

import oracle.port as P
from conftest import golden_names, load_golden
from reveal_b200 import synth
from util import check_handle_reuse, check_pack_block, check_sweep_prefetch, NativeIndex, assert_same, check_against_golden, check_against_oracle, random_related

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_cuda_matches_golden(cuda_lib, name):
    check_against_golden(cuda_lib, load_golden(name))


@pytest.mark.parametrize("nsamples,length,sigma,minl", [
    (2, 1500, 2, 4), (3, 3000, 4, 6), (2, 9000, 4, 8), (5, 2500, 3, 5),
    (2, 300000, 4, 12), (4, 200000, 4, 12), (7, 50000, 4, 10), (2, 40000, 2, 16),
])
def test_cuda_matches_oracle_random(cuda_lib, nsamples, length, sigma, minl):
    rng = np.random.default_rng(nsamples * 1000 + length)
    T, nsep, _ = P.assemble(random_related(rng, nsamples, length, sigma))
    check_against_oracle(cuda_lib, T, nsep, nsamples, minl=minl)


def test_cuda_tiny_and_single_sample(cuda_lib):
    for text in (b"A", b"AC", b"ACA", b"GATTACA", b"A" * 5000, b"AC" * 3000):
        T, nsep, _ = P.assemble([[text]])
        check_against_oracle(cuda_lib, T, nsep, 1, minl=0)


def test_cuda_ragged_many_contigs(cuda_lib):
    """Many '$' (graph inputs put one per segment), empty-ish and 1-bp contigs."""
    rng = np.random.default_rng(11)
    al = np.frombuffer(b"ACGT", np.uint8)
    base = al[rng.integers(0, 4, size=60000)].tobytes()
    cuts = np.sort(rng.integers(0, len(base), size=400))
    s0 = [base[a:b] for a, b in zip([0] + cuts.tolist(), cuts.tolist() + [len(base)]) if b > a]
    mut = bytearray(base)
    for p in rng.integers(0, len(base), size=600):
        mut[p] = al[rng.integers(0, 4)]
    cuts = np.sort(rng.integers(0, len(base), size=300))
    s1 = [bytes(mut[a:b]) for a, b in zip([0] + cuts.tolist(), cuts.tolist() + [len(base)]) if b > a] + [b"A", b"C"]
    T, nsep, _ = P.assemble([s0, s1])
    check_against_oracle(cuda_lib, T, nsep, 2, minl=10)


def test_cuda_rc_and_alphabets(cuda_lib):
    rng = np.random.default_rng(7)
    T, nsep, _ = P.assemble(random_related(rng, 2, 50000, 4))
    check_against_oracle(cuda_lib, T, nsep, 2, rc=1, minl=10)
    # IUPAC + N + lowercase (the reference keeps case with --noupper)
    T, nsep, _ = P.assemble(random_related(rng, 3, 30000, 16, alphabet=b"ACGTNRYKMSWBDHVn"))
    check_against_oracle(cuda_lib, T, nsep, 3, minl=4)
    T, nsep, _ = P.assemble(random_related(rng, 2, 30000, 8, alphabet=b"ACGTacgt"))
    check_against_oracle(cuda_lib, T, nsep, 2, minl=6)


@pytest.mark.parametrize("ngenomes,length", [(2, 500000), (3, 300000)])
def test_cuda_matches_oracle_synthetic_genomes(cuda_lib, ngenomes, length):
    """The bench recipe (1% SNP, 0.1% indel) at a size the oracle does in a second."""
    T, nsep, ns = synth.workload(ngenomes, length, seed=3)
    check_against_oracle(cuda_lib, T, nsep, ns, minl=20)


def _properties(T, nsep, ns, idx, minl):
    n = len(T)
    SA, SAi, LCP = idx.arr("SA"), idx.arr("SAi"), idx.arr("LCP")
    # permutation + inverse
    assert np.array_equal(SAi[SA], np.arange(n, dtype=np.int32))
    # sortedness on a sample of adjacent pairs, and LCP against a direct byte comparison
    rng = np.random.default_rng(0)
    Tb = T.tobytes()
    for r in rng.integers(1, n, size=4000).tolist():
        a, b = int(SA[r - 1]), int(SA[r])
        h = 0
        while a + h < n and b + h < n and Tb[a + h] == Tb[b + h]:
            h += 1
        assert a + h == n or (b + h < n and Tb[a + h] < Tb[b + h]), "SA order violated at rank %d" % r
        stop = h
        for k in range(h):  # barrier rule of compute_lcp (interface.c:107)
            if Tb[b + k] in (36, 78):
                stop = k
                break
        assert int(LCP[r]) == stop, "LCP mismatch at rank %d" % r
    # every reported MUM is an exact match, left- and right-maximal, one position per sample
    mums = idx.mums(minl)
    assert len(mums) > 0
    for l, a, b in mums[:: max(1, len(mums) // 3000)].tolist():
        assert Tb[a:a + l] == Tb[b:b + l]
        assert a <= nsep[0] < b
        assert a == 0 or Tb[a - 1] != Tb[b - 1] or Tb[a - 1] in (36, 78)
        assert Tb[a + l] != Tb[b + l] or Tb[a + l] in (36, 78)
    return mums


def test_cuda_full_size_c2(cuda_lib):
    """BASELINE configs[1]: 2 synthetic 5 Mbp genomes -- properties, then bit-exact vs the oracle."""
    T, nsep, ns = synth.workload(2, 5_000_000, seed=1)
    with NativeIndex(cuda_lib, T, nsep, ns) as idx:
        mums = _properties(T, nsep, ns, idx, 20)
        o = P.Index(T, nsep, ns)
        assert_same(idx.arr("SA"), o.SA, "SA")
        assert_same(idx.arr("LCP"), o.LCP, "LCP")
        assert_same(mums, o.getmums(20), "getmums")
        # idempotence: a second build on the same handle (workspace reuse) gives the same answer
    with NativeIndex(cuda_lib, T, nsep, ns) as idx2:
        assert_same(idx2.mums(20), mums, "getmums (rebuild)")


def test_cuda_full_size_c3_multi(cuda_lib):
    """BASELINE configs[2]: 5 synthetic 5 Mbp genomes, multi-MUM sweep vs the oracle."""
    T, nsep, ns = synth.workload(5, 5_000_000, seed=1)
    o = P.Index(T, nsep, ns)
    with NativeIndex(cuda_lib, T, nsep, ns) as idx:
        assert_same(idx.arr("SA"), o.SA, "SA")
        assert_same(idx.arr("LCP"), o.LCP, "LCP")
        assert_same(idx.arr("SO"), o.SO, "SO")
        hdr, mem = idx.multimums(20, 2)
        oh, om = o.getmultimums(20, 2)
        assert_same(hdr, oh, "getmultimums hdr")
        assert_same(mem, om, "getmultimums members")


def test_reveallib_surface_on_gpu():
    """The drop-in class end to end (reference test01 input, test_reveal.py:36-41)."""
    from reveal_b200 import reveallib, reveallib64
    for mod in (reveallib, reveallib64):
        idx = mod.index()
        idx.addsample("1")
        assert idx.addsequence("ACTTGCTAGCTAGTCAG") == (0, 17)
        idx.addsample("2")
        assert idx.addsequence("ACTAGCTAGCTAGTGAG") == (18, 35)
        idx.construct()
        g = load_golden("test01_pair")
        assert idx.SA == g["SA"].tolist() and idx.LCP == g["LCP"].tolist() and idx.SAi == g["SAi"].tolist()
        assert idx.getmums(1) == [(l, (a, b), 0) for l, a, b in g["mums"].tolist()]
        assert idx.nsep == [17] and idx.n == 36 and idx.nsamples == 2


@pytest.mark.parametrize("bits", ["32", "64"])
def test_cuda_both_key_widths(cuda_lib, monkeypatch, bits):
    """RV_SA_KEY_BITS forces the 32- / 64-bit k-mer key path of the SA builder."""
    monkeypatch.setenv("RV_SA_KEY_BITS", bits)
    rng = np.random.default_rng(21)
    T, nsep, _ = P.assemble(random_related(rng, 3, 120000, 4))
    check_against_oracle(cuda_lib, T, nsep, 3, minl=10)
    for name in ("tandem", "all_A", "iupac_mix", "with_N_d2"):
        check_against_golden(cuda_lib, load_golden(name))
    T, nsep, ns = synth.workload(2, 400000, seed=5)
    check_against_oracle(cuda_lib, T, nsep, ns, minl=20)


def test_cuda_long_identical_and_repeats(cuda_lib):
    """Comparisons longer than the cap and groups larger than SA_SMALL_G go through the doubling rounds + Kasai."""
    rng = np.random.default_rng(3)
    al = np.frombuffer(b"ACGT", np.uint8)
    s = al[rng.integers(0, 4, size=70000)].tobytes()
    t = bytearray(s)
    t[65000] = ord("A") if t[65000] != ord("A") else ord("C")
    T, nsep, _ = P.assemble([[s], [bytes(t)], [s[100:69000]]])
    check_against_oracle(cuda_lib, T, nsep, 3, minl=10)
    rep = al[rng.integers(0, 4, size=60)].tobytes()
    s0 = al[rng.integers(0, 4, size=15000)].tobytes() + rep * 40 + al[rng.integers(0, 4, size=15000)].tobytes()
    s1 = s0[:7000] + b"G" + s0[7001:20000] + rep * 3 + s0[20000:]
    T, nsep, _ = P.assemble([[s0], [s1]])
    check_against_oracle(cuda_lib, T, nsep, 2, minl=8)


def test_cuda_byte_comparison_path(cuda_lib, monkeypatch):
    """RV_SA_NO_PACK forces the byte-wise comparison loops that large alphabets use."""
    monkeypatch.setenv("RV_SA_NO_PACK", "1")
    T, nsep, ns = synth.workload(3, 300000, seed=8)
    check_against_oracle(cuda_lib, T, nsep, ns, minl=20)
    check_against_golden(cuda_lib, load_golden("with_N_d2"))


def test_cuda_handle_reuse_alphabet_cache(cuda_lib):
    check_handle_reuse(cuda_lib)


def test_cuda_sweeps_between_rebuilds(cuda_lib):
    check_sweep_prefetch(cuda_lib)


def test_cuda_pack_into_peer_block(cuda_lib):
    check_pack_block(cuda_lib)


def test_cuda_anchor_units_single_rank(cuda_lib):
    """shard.anchor_units on one GPU (no process group): every unit's rows equal the oracle's."""
    from reveal_b200 import shard
    rng = np.random.default_rng(12)
    units = []
    for ns, length in ((2, 30000), (3, 12000), (2, 5000)):
        T, nsep, _ = P.assemble(random_related(rng, ns, length, 4))
        units.append((T, np.asarray(nsep, dtype=np.int64), ns))
    got = shard.anchor_units(units, minl=10, lib=cuda_lib)
    for (T, nsep, ns), res in zip(units, got):
        o = P.Index(T, nsep, ns)
        if ns == 2:
            assert_same(res, o.getmums(10, rem=True), "unit rows")
        else:
            oh, om = o.getmultimums(10, 2)
            assert_same(res[0], oh, "unit hdr rows")
            assert_same(res[1], om, "unit member rows")


def test_cuda_full_size_c4_properties(cuda_lib):
    """BASELINE configs[3] size (2 x 100 Mbp, n = 2e8, 64-bit k-mer keys): size-independent properties only."""
    T, nsep, ns = synth.workload(2, 100_000_000, seed=1)
    with NativeIndex(cuda_lib, T, nsep, ns) as idx:
        mums = _properties(T, nsep, ns, idx, 20)
        assert len(mums) > 500000


def test_cuda_getmultimems_matches_oracle(cuda_lib):
    """getmultimems through the drop-in class, incl. the reference's `continue` quirk, vs the oracle."""
    from reveal_b200 import reveallib
    rng = np.random.default_rng(99)
    al = np.frombuffer(b"ACGT", np.uint8)
    base = al[rng.integers(0, 4, size=30000)].tobytes()
    rep = base[500:900]
    raw = [base + rep + base[:1000], base[:20000] + b"T" + base[20001:] + rep, rep + base[1000:25000] + rep + rep, base[5000:28000]]
    idx = reveallib.index()
    for k, s in enumerate(raw):
        idx.addsample("s%d" % k)
        idx.addsequence(s.decode())
    idx.construct()
    T = np.frombuffer(idx.T.encode(), np.uint8)
    o = P.Index(T, idx.nsep, 4)
    for minl, minn in ((20, 2), (12, 3), (30, 4)):
        got = idx.getmultimems(minl, minn)
        assert got == P.multi_to_tuples(*o.getmultimems(minl, minn))
        assert len(got) > 0


# ---- repeat-bearing inputs at Mbp scale (the doubling rounds and the LCP fallback run here) -----------------------------------
@pytest.mark.parametrize("tag,expect", [("2", 6427), ("3", 17596), ("123", 24596)])
def test_cuda_real_data_matches_reference(cuda_lib, tag, expect):
    """The reference's own repeat-bearing fixtures (2a/2b, 3a/3b, 123a/123b; n up to 10 754 553): SA / SAi / LCP digests and the
    getmums(20) list of the UNMODIFIED reference (tests/golden/make_real_golden.py), known-answer counts of BASELINE.md section 2."""
    from util import check_against_real
    info = {}

    def build(T, nsep):
        with NativeIndex(cuda_lib, T, nsep, 2) as idx:
            info.update(idx.times())
            return idx.arr("SA"), idx.arr("SAi"), idx.arr("LCP"), idx.mums(20)

    assert check_against_real(build, tag) == expect
    print("real %s: %s" % (tag, info))


def test_cuda_repeat_genomes_full_size(cuda_lib):
    """2 x 5 Mbp on a repeat-bearing ancestor (families, tandem arrays, segmental duplications, N runs) vs the oracle."""
    T, nsep, ns = synth.repeat_workload(2, 5_000_000, seed=1)
    o = P.Index(T, nsep, ns)
    with NativeIndex(cuda_lib, T, nsep, ns) as idx:
        assert idx.times()["sa_rounds"] > 0, "this input is meant to need the doubling rounds"
        assert_same(idx.arr("SA"), o.SA, "SA")
        assert_same(idx.arr("SAi"), o.SAi, "SAi")
        assert_same(idx.arr("LCP"), o.LCP, "LCP")
        assert_same(idx.mums(20, 1), o.getmums(20, rem=True), "getmums_rem")
        hdr, mem = idx.multimums(20, 2)
        oh, om = o.getmultimums(20, 2)
        assert_same(hdr, oh, "getmultimums hdr")
        assert_same(mem, om, "getmultimums members")


def test_cuda_repeat_genomes_three_samples(cuda_lib):
    T, nsep, ns = synth.repeat_workload(3, 1_500_000, seed=2)
    check_against_oracle(cuda_lib, T, nsep, ns, minl=20)


def test_cuda_graph_like_text(cuda_lib):
    """One '$' per node: 2 x 2.5e5 contigs of ~20 bp (the text shape of graph input), n = 1.05e7."""
    T, nsep, ns = synth.graph_like_workload(250_000, 20, seed=3)
    check_against_oracle(cuda_lib, T, nsep, ns, minl=12)


def test_cuda_highly_repetitive(cuda_lib):
    """Kasai-fallback territory: most suffixes go through the doubling rounds (long tandem arrays, many identical copies)."""
    rng = np.random.default_rng(17)
    al = np.frombuffer(b"ACGT", np.uint8)
    unit = al[rng.integers(0, 4, size=3000)].tobytes()
    s0 = unit * 120 + al[rng.integers(0, 4, size=50000)].tobytes() + (b"ACGTTGCA" * 20000)
    s1 = bytearray(s0)
    for p in rng.integers(0, len(s1), size=300):
        s1[p] = al[rng.integers(0, 4)]
    T, nsep, _ = P.assemble([[s0], [bytes(s1)]])
    check_against_oracle(cuda_lib, T, nsep, 2, minl=20)


def test_cuda_sweeps_over_caller_supplied_device_arrays(cuda_lib):
    """rv_sweep_pair_device / rv_sweep_multi_device: the sweeps over device arrays a caller holds (here: the arrays of a built index,
    through rv_device_arrays) give the rows of the index's own sweeps."""
    import ctypes

    from reveal_b200 import _native
    L = cuda_lib
    rng = np.random.default_rng(31)
    for ns, length in ((2, 60000), (4, 25000)):
        T, nsep, _ = P.assemble(random_related(rng, ns, length, 4))
        with NativeIndex(L, T, nsep, ns) as idx:
            ptr = [ctypes.c_void_p() for _ in range(5)]
            _native.check(L, L.rv_device_arrays(idx.h, *[ctypes.byref(p) for p in ptr]))
            dT, dSA, dSAi, dLCP, dSO = ptr
            assert dT.value and dSA.value and dLCP.value and (dSO.value or ns == 2)
            ws = ctypes.c_void_p()
            _native.check(L, L.rv_index_create(ctypes.byref(ws), None))   # a second handle as the workspace of the foreign arrays
            try:
                if ns == 2:
                    want = idx.mums(12, 1)
                    c = ctypes.c_int64()
                    _native.check(L, L.rv_sweep_pair_device(ws, dT, dSA, dLCP, idx.n, idx.n, int(nsep[0]), 0, 1, 12, ctypes.byref(c)))
                    rows = np.empty((c.value, 3), np.int64)
                    _native.check(L, L.rv_mums_pair_fetch(ws, rows.ctypes.data, c.value))
                    assert_same(rows, want, "rv_sweep_pair_device")
                    assert len(rows) > 0
                wh, wm = idx.multimums(12, 2)
                nr, nm = ctypes.c_int64(), ctypes.c_int64()
                _native.check(L, L.rv_sweep_multi_device(ws, dT, dSA, dLCP, dSO, idx.n, int(nsep[0]), ns, 12, 2, ctypes.byref(nr), ctypes.byref(nm)))
                hdr = np.empty((nr.value, 3), np.int64)
                mem = np.empty((nm.value, 2), np.int64)
                _native.check(L, L.rv_mums_multi_fetch(ws, hdr.ctypes.data, nr.value, mem.ctypes.data, nm.value))
                assert_same(hdr, wh, "rv_sweep_multi_device hdr")
                assert_same(mem, wm, "rv_sweep_multi_device members")
                assert len(hdr) > 0
            finally:
                L.rv_index_free(ws)
