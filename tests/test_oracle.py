"""Pins the CPU oracle (oracle/reveal_oracle.c via oracle.port): against the golden
vectors minted from the unmodified reference build, and -- when the reference
tree is present (build container) -- against that build directly on the
reference's own FASTA fixtures."""
import os

import numpy as np
import pytest

import oracle.port as P
import oracle.ref as R
from conftest import golden_names, load_golden
from util import assert_same


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    g = load_golden(name)
    ns = int(g["nsamples"])
    o = P.Index(g["T_in"], g["nsep"], ns, int(g["rc"]))
    assert_same(o.T, g["T_indexed"], "T")
    assert_same(o.SA, g["SA"], "SA")
    assert_same(o.SAi, g["SAi"], "SAi")
    assert_same(o.LCP, g["LCP"], "LCP")
    assert_same(o.getmums(int(g["minl"])), g["mums"], "getmums")
    if ns > 2:
        assert_same(o.SO, g["SO"], "SO")
        hdr, mem = o.getmultimums(int(g["minl"]), int(g["minn"]))
        assert_same(hdr, g["mm_hdr"], "getmultimums hdr")
        assert_same(mem, g["mm_mem"], "getmultimums members")


def test_comp_table_matches_reference_rule():
    t = P.comp_table()
    assert bytes(t[[ord(c) for c in "ACGTUNRYKMBVDHSW"]]) == b"TGCAANYRMKVBHDSW"
    assert t[96] == 64 and all(t[c] == c for c in range(64))


def _fasta(path):
    seqs, cur = [], []
    for line in open(path):
        line = line.strip()
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur))
            cur = []
        elif line:
            cur.append(line.upper())
    if cur:
        seqs.append("".join(cur))
    return seqs


REF_TESTS = os.path.join(R.REFERENCE_ROOT, "tests")
needs_ref = pytest.mark.skipif(not (os.path.isdir(REF_TESTS) and R.available()), reason="reference tree / oracle/_ref not present")


@needs_ref
@pytest.mark.parametrize("files,minl,expect", [(("1a.fa", "1b.fa"), 20, 553), (("1c.fa", "1d.fa"), 20, 10)])
def test_oracle_vs_reference_pair(files, minl, expect):
    """Full-size pair fixtures; MUM counts are the known answers of SURVEY.md section 6."""
    samples = [_fasta(os.path.join(REF_TESTS, f)) for f in files]
    ridx = R.index_from_samples(samples)
    T, nsep, _ = P.assemble(samples)
    o = P.Index(T, nsep, len(samples))
    assert_same(o.SA, np.asarray(ridx.SA, np.int32), "SA")
    assert_same(o.LCP, np.asarray(ridx.LCP, np.int32), "LCP")
    ref = np.asarray([(l, a, b) for l, (a, b), _ in ridx.getmums(minl)], np.int64).reshape(-1, 3)
    assert len(ref) == expect
    assert_same(o.getmums(minl), ref, "getmums")


@needs_ref
def test_oracle_vs_reference_multi():
    samples = [_fasta(os.path.join(REF_TESTS, f)) for f in ("1a.fa", "1b.fa", "1c.fa")]
    ridx = R.index_from_samples(samples)
    T, nsep, _ = P.assemble(samples)
    o = P.Index(T, nsep, 3)
    assert_same(o.SO, np.asarray(ridx.SO, np.uint16), "SO")
    mm = ridx.getmultimums(minlength=20, minn=2)
    assert len(mm) == 558
    assert P.multi_to_tuples(*o.getmultimums(20, 2)) == [tuple(x) for x in mm]


@pytest.mark.parametrize("tag,expect", [("1", 553), ("2", 6427)])
def test_oracle_matches_reference_on_real_data(tag, expect):
    """The reference's real-data fixtures (repeat-bearing Aspergillus niger contigs) from the committed
    compact copy: the restatement reproduces the unmodified reference's SA / SAi / LCP digests and its getmums(20) list."""
    from util import check_against_real

    def build(T, nsep):
        o = P.Index(T, nsep, 2)
        return o.SA, o.SAi, o.LCP, o.getmums(20)

    assert check_against_real(build, tag) == expect
