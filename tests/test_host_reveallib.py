"""Host-side logic of the drop-in `reveallib.index` class against the reference's own extension
(oracle/_ref, unmodified interface.c): same return values, getters and error behaviour.  The device
work runs on the emulated kernels (CPU tier); the gpu-marked twin of this file is test_gpu_parity.py."""
import numpy as np
import pytest

import oracle.ref as R
from conftest import load_golden

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref (compiled reference) not present")

SAMPLES = [["ACGTACGTTAGCATGCAGGATCCA", "TTGACA"], ["ACGTACCTTAGCATGCAGGTTCCA"], ["GGATCCAACGTACGTTAG", "A", "CCGT"]]


def build_both(reveallib, samples, rc=0, bits=32):
    ours, ref = (reveallib.index() if bits == 32 else reveallib.index64()), R.module(bits).index()
    rets = []
    for k, seqs in enumerate(samples):
        rets.append((ours.addsample("s%d" % k), ref.addsample("s%d" % k)))
        for s in seqs:
            rets.append((ours.addsequence(s), ref.addsequence(s)))
    for a, b in rets:
        assert a == b
    if rc:
        ours.construct(rc=1)
        ref.construct(rc=1)
    else:
        ours.construct()
        ref.construct()
    return ours, ref


@needs_ref
@pytest.mark.parametrize("bits", [32, 64])
def test_surface_matches_reference(emu_reveallib, bits):
    ours, ref = build_both(emu_reveallib, SAMPLES, bits=bits)
    for attr in ("n", "nsamples", "samples", "nsep", "nodes", "depth", "leftnode", "rightnode", "SA", "SAi", "LCP", "SO"):
        assert getattr(ours, attr) == getattr(ref, attr), attr
    assert ours.T == ref.T[:ref.n]
    assert ours.getmultimums(minlength=3, minn=2) == [tuple(m) for m in ref.getmultimums(minlength=3, minn=2)]
    assert ours.getmultimums(3, 3) == [tuple(m) for m in ref.getmultimums(minlength=3, minn=3)]


@needs_ref
def test_pair_getmums_and_rc(emu_reveallib):
    ours, ref = build_both(emu_reveallib, SAMPLES[:2])
    assert ours.getmums(3) == [tuple(m) for m in ref.getmums(3)]
    with pytest.raises(TypeError):
        ours.SO  # "SO not available." for two samples (interface.c:575-579)
    with pytest.raises(TypeError):
        ref.SO
    ours, ref = build_both(emu_reveallib, SAMPLES[:2], rc=1)
    assert ours.T == ref.T[:ref.n]  # the second sample is reverse-complemented in place (interface.c:168-172)
    assert ours.getmums(3) == [tuple(m) for m in ref.getmums(3)]
    assert ours.SA == ref.SA and ours.LCP == ref.LCP


@needs_ref
def test_error_behaviour(emu_reveallib):
    ours, ref = emu_reveallib.index(), R.module(32).index()
    with pytest.raises(emu_reveallib.error):
        ours.construct()  # "No text to index."
    with pytest.raises(R.module(32).error):
        ref.construct()
    with pytest.raises(emu_reveallib.error):
        ours.addsample(3)  # "Sample name has to be a string."
    with pytest.raises(R.module(32).error):
        ref.addsample(3)
    with pytest.raises(TypeError):
        ours.SA  # "Index not yet constructed."
    with pytest.raises(TypeError):
        ref.SA
    with pytest.raises(emu_reveallib.error):
        ours.align(lambda *a, **k: (), lambda *a: None)  # "Index not yet constructed, alignment stopped."


def test_empty_sample_and_one_bp_contigs(emu_reveallib):
    """nsep may repeat (a sample without sequences); 1-bp contigs (reference fixtures t1.fa / t2.fa)."""
    import oracle.port as P
    idx = emu_reveallib.index()
    idx.addsample("a")
    idx.addsequence("ACGTTGCA")
    idx.addsample("empty")
    idx.addsample("c")
    for s in ("A", "C", "ACGTAGCA"):
        idx.addsequence(s)
    idx.construct()
    assert idx.nsep == [8, 8]
    T = np.frombuffer(idx.T.encode(), np.uint8)
    o = P.Index(T, idx.nsep, 3)
    assert idx.SA == o.SA.tolist() and idx.LCP == o.LCP.tolist() and idx.SO == o.SO.tolist()
    assert idx.getmultimums(2, 2) == P.multi_to_tuples(*o.getmultimums(2, 2))


def test_golden_through_the_class(emu_reveallib):
    g = load_golden("1a_1b_1c_triple")
    idx = emu_reveallib.index()
    T = g["T_in"].tobytes().decode()
    bounds = [0] + [int(x) + 1 for x in g["nsep"]] + [len(T)]
    for k in range(3):
        idx.addsample("s%d" % k)
        idx.addsequence(T[bounds[k]:bounds[k + 1] - 1])
    idx.construct()
    assert idx.SA == g["SA"].tolist() and idx.LCP == g["LCP"].tolist()
    got = idx.getmultimums(int(g["minl"]), int(g["minn"]))
    want = [(l, n, tuple(map(tuple, g["mm_mem"][f:f + n].tolist()))) for l, n, f in g["mm_hdr"].tolist()]
    assert got == want
    t = idx.times()
    assert t["launches"] > 0


def test_cache_files_round_trip(emu_reveallib, tmp_path, monkeypatch):
    """cache=1 writes .reveal.t/.sa/.lcp (interface.c:273-285); index(sa=, lcp=) reads them back (:224-231,255-262)."""
    monkeypatch.chdir(tmp_path)
    cache_round_trip(emu_reveallib, tmp_path)


@pytest.mark.gpu
def test_cache_files_round_trip_cuda(tmp_path, monkeypatch):
    """The same on the CUDA library: rv_build_cached with and without the LCP file (isa_scatter / isa_verify kernels, the
    sparse Kasai kernel with every position marked), and a suffix array file that is no permutation."""
    from reveal_b200 import reveallib
    monkeypatch.chdir(tmp_path)
    cache_round_trip(reveallib, tmp_path)


def cache_round_trip(emu_reveallib, tmp_path):
    a = emu_reveallib.index(cache=1)
    for k, seqs in enumerate(SAMPLES):
        a.addsample("s%d" % k)
        for s in seqs:
            a.addsequence(s)
    a.construct()
    assert (tmp_path / ".reveal.sa").stat().st_size == 4 * a.n
    for kw in ({"sa": ".reveal.sa", "lcp": ".reveal.lcp"}, {"sa": ".reveal.sa"}):
        b = emu_reveallib.index(**kw)
        for k, seqs in enumerate(SAMPLES):
            b.addsample("s%d" % k)
            for s in seqs:
                b.addsequence(s)
        b.construct()
        assert b.SA == a.SA and b.SAi == a.SAi and b.LCP == a.LCP and b.SO == a.SO
        assert b.getmultimums(3, 2) == a.getmultimums(3, 2)
    bad = np.zeros(a.n, np.int32)
    bad.tofile("bad.sa")
    c = emu_reveallib.index(sa="bad.sa")
    for k, seqs in enumerate(SAMPLES):
        c.addsample("s%d" % k)
        for s in seqs:
            c.addsequence(s)
    with pytest.raises(emu_reveallib.error):
        c.construct()


@needs_ref
@pytest.mark.parametrize("minl,minn", [(3, 2), (2, 3), (6, 2), (1, 2), (0, 2)])
def test_getmultimems_matches_reference(emu_reveallib, minl, minn):
    """getmultimems incl. the reference's `continue` quirk (reveal.c:340-342), against the unmodified extension."""
    rng = np.random.default_rng(minl * 10 + minn)
    al = np.frombuffer(b"ACGT", np.uint8)
    base = al[rng.integers(0, 4, size=400)].tobytes().decode()
    rep = base[50:90]
    samples = [[base + rep + base[:100]], [base[:200] + "T" + base[201:] + rep], [rep + base[100:350] + rep + rep]]
    ours, ref = build_both(emu_reveallib, samples)
    want = [tuple(m) for m in ref.getmultimems(minlength=minl, minn=minn)]
    got = ours.getmultimems(minl, minn)
    assert got == want
    if minl >= 2:
        assert len(want) > 0


def test_getmultimems_matches_oracle_port(emu_reveallib):
    import oracle.port as P
    from util import random_related
    rng = np.random.default_rng(99)
    samples = random_related(rng, 4, 1500, 3, snp=0.05)
    idx = emu_reveallib.index()
    for k, seqs in enumerate(samples):
        idx.addsample("s%d" % k)
        for s in seqs:
            idx.addsequence(s.decode())
    idx.construct()
    T = np.frombuffer(idx.T.encode(), np.uint8)
    o = P.Index(T, idx.nsep, 4)
    for minl, minn in ((5, 2), (4, 3), (8, 4)):
        assert idx.getmultimems(minl, minn) == P.multi_to_tuples(*o.getmultimems(minl, minn))


def test_copy_is_independent(emu_reveallib):
    a = emu_reveallib.index()
    for k, seqs in enumerate(SAMPLES):
        a.addsample("s%d" % k)
        for s in seqs:
            a.addsequence(s)
    a.construct()
    b = a.copy()
    assert b.SA == a.SA and b.SAi == a.SAi and b.LCP == a.LCP and b.SO == a.SO and b.T == a.T and b.nsep == a.nsep
    assert b.getmultimums(3, 2) == a.getmultimums(3, 2)
    del a
    assert len(b.getmultimums(3, 2)) > 0


def test_addsequence_copies_big_sequences_in_parts(emu_reveallib):
    """A sequence of 2 MB and more is copied into the text buffer on several threads (copy_big): the text that comes back is the
    text that went in, whatever the lengths are relative to the part size, and the returned intervals are the reference's."""
    rng = np.random.default_rng(9)
    al = np.frombuffer(b"ACGTN", np.uint8)
    idx = emu_reveallib.index()
    want, at = [], 0
    for k, ln in enumerate([(1 << 21) - 1, 1 << 21, (1 << 21) + 4097, 3 * (1 << 20) + 11, 5]):
        s = al[rng.integers(0, 5, size=ln)].tobytes().decode()
        idx.addsample("s%d" % k)
        assert idx.addsequence(s) == (at, at + ln)
        want.append(s)
        at += ln + 1
    assert idx.n == at
    assert idx.T == "$".join(want) + "$"
