"""Mints the golden vectors in tests/golden/*.npz from the UNMODIFIED reference
extension (oracle/_ref, built from /root/reference by oracle/ref/Makefile).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py

Every case stores its INPUT (samples as lists of sequences -> T, nsep) next to
the reference's OUTPUT (SA, SAi, LCP, SO, getmums, getmultimums), so the tests
never need /root/reference at run time.  Inputs are the reference's own test
inputs (reveal/tests/test_reveal.py:36-41 and slices of tests/*.fa) plus small
adversarial texts; slices keep the files small.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle.ref as ref  # noqa: E402

REF_TESTS = os.path.join(ref.REFERENCE_ROOT, "tests")


def fasta(name, upper=True):
    """Sequences of a reference fixture, upper-cased like utils.fasta_reader's default (utils.py:79-144)."""
    seqs, cur = [], []
    with open(os.path.join(REF_TESTS, name)) as f:
        for line in f:
            line = line.strip()
            if line.startswith(">"):
                if cur:
                    seqs.append("".join(cur))
                cur = []
            elif line:
                cur.append(line.upper() if upper else line)
    if cur:
        seqs.append("".join(cur))
    return seqs


def cases():
    a, b, c = fasta("1a.fa")[0], fasta("1b.fa")[0], fasta("1c.fa")[0]
    brc = fasta("1brc.fa")[0]
    d_low = fasta("1d.fa", upper=False)[0]
    d2 = fasta("d2.fa")[0]
    e = fasta("1e.fa")
    npos = d2.find("N")
    out = []
    # the reference's own in-memory pair (test_reveal.py:36-41, minlength=1)
    out.append(("test01_pair", [["ACTTGCTAGCTAGTCAG"], ["ACTAGCTAGCTAGTGAG"]], 0, 1, 2))
    out.append(("t1_t2_one_bp", [fasta("t1.fa"), fasta("t2.fa")], 0, 1, 2))
    out.append(("1a_1b_head3k", [[a[:3000]], [b[:3000]]], 0, 10, 2))
    out.append(("1a_1b_mid20k", [[a[40000:60000]], [b[30000:50000]]], 0, 20, 2))
    out.append(("1a_1brc_rc", [[a[:4000]], [brc[-4000:]]], 1, 12, 2))
    out.append(("1a_1b_1c_triple", [[a[:2500]], [b[:2500]], [c[:2500]]], 0, 8, 2))
    out.append(("1a_1b_1c_minn3", [[a[10000:16000]], [b[9000:15000]], [c[10000:16000]]], 0, 15, 3))
    out.append(("lowercase_1d", [[d_low[:3000]], [a[:3000]]], 0, 8, 2))
    out.append(("with_N_d2", [[d2[max(0, npos - 1500):npos + 1500]], [b[:3000]]], 0, 6, 2))
    out.append(("multicontig_1e", [[x[:1500] for x in e], [b[:3000]]], 0, 8, 2))
    out.append(("five_samples", [[a[:1500]], [b[:1500]], [c[:1500]], [a[200:1700]], [c[100:1600]]], 0, 6, 2))
    out.append(("all_A", [["A" * 700], ["A" * 500]], 0, 2, 2))
    out.append(("iupac_mix", [["ACGTRYKMSWBDHVNACGTNNNNACGTACGTRYACGT" * 9], ["ACGTRYKMSWBDHVNACGAACGTACGTRYACGT" * 10]], 0, 4, 2))
    out.append(("tandem", [["ACGACGACGACGACGACGTTT" * 40], ["ACGACGACGACGACGTTTACG" * 40], ["GACGACGACGTTTACGACG" * 30]], 0, 5, 2))
    return out


def mint(name, samples, rc, minl, minn):
    idx = ref.index_from_samples(samples, bits=32, rc=rc)
    n = idx.n
    T = np.frombuffer(idx.T.encode("latin-1"), dtype=np.uint8)[:n]
    # T as it was BEFORE the reference's in-place rc is what a caller passes in
    idx0 = ref.index_from_samples(samples, bits=32, construct=False)
    T0 = np.frombuffer(idx0.T.encode("latin-1"), dtype=np.uint8)[:n]
    nsamples = idx.nsamples
    mums = np.asarray([(l, ab[0], ab[1]) for l, ab, _ in idx.getmums(minl)], dtype=np.int64).reshape(-1, 3)
    d = dict(T_in=T0, T_indexed=T, nsep=np.asarray(idx.nsep, dtype=np.int64), nsamples=np.int32(nsamples), rc=np.int32(rc),
             minl=np.int32(minl), minn=np.int32(minn), SA=np.asarray(idx.SA, dtype=np.int32), SAi=np.asarray(idx.SAi, dtype=np.int32),
             LCP=np.asarray(idx.LCP, dtype=np.int32), mums=mums)
    if nsamples > 2:
        d["SO"] = np.asarray(idx.SO, dtype=np.uint16)
        mm = idx.getmultimums(minlength=minl, minn=minn)
        hdr, mem = [], []
        for l, cnt, members in mm:
            hdr.append((l, cnt, len(mem)))
            mem.extend(members)
        d["mm_hdr"] = np.asarray(hdr, dtype=np.int64).reshape(-1, 3)
        d["mm_mem"] = np.asarray(mem, dtype=np.int64).reshape(-1, 2)
    # 64-bit build of the reference must agree (reveallib64)
    idx64 = ref.index_from_samples(samples, bits=64, rc=rc)
    assert list(idx64.SA) == list(idx.SA) and list(idx64.LCP) == list(idx.LCP)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print("%-20s n=%-6d mums=%-5d multi=%s" % (name, n, len(mums), len(d.get("mm_hdr", []))))


if __name__ == "__main__":
    if not ref.build():
        sys.exit("reference sources not present")
    for case in cases():
        mint(*case)
