"""Mints tests/golden/real/asp_niger.npz: the reference's repeat-bearing real-data fixtures
(tests/123a.fa + tests/123b.fa, three Aspergillus niger contigs each; contig 2 / 3 of each file are
tests/2a.fa,2b.fa / 3a.fa,3b.fa) in a compact form, next to what the UNMODIFIED reference extension
(oracle/_ref) computes on them.

Run in the build container only (needs /root/reference):
    python tests/golden/make_real_golden.py

Stored: the six contigs as 2-bit codes (ACGT) plus an exception list for the few IUPAC letters, and per
pair (1a/1b = contig 0, 2a/2b = contig 1, 3a/3b = contig 2, 123a/123b = all three) the reference's answers:
n, sha256 of the SA / SAi / LCP arrays (little-endian int32), and the full getmums(20) list -- the
known-answer counts of BASELINE.md section 2 (553 / 6 427 / 17 596 / 24 596) fall out of it.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import oracle.ref as ref  # noqa: E402
from make_golden import fasta  # noqa: E402

CODE = np.full(256, 255, np.uint8)
for k, c in enumerate(b"ACGT"):
    CODE[c] = k


def pack(seq):
    a = np.frombuffer(seq.encode("ascii"), np.uint8)
    c = CODE[a]
    exc_pos = np.flatnonzero(c == 255).astype(np.int64)
    exc_chr = a[exc_pos]
    c = np.where(c == 255, 0, c).astype(np.uint8)
    pad = (-len(c)) % 4
    c = np.concatenate([c, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    packed = (c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)).astype(np.uint8)
    return packed, len(a), exc_pos, exc_chr


def digest(arr, dtype=np.int32):
    return hashlib.sha256(np.ascontiguousarray(np.asarray(arr, dtype=dtype)).tobytes()).hexdigest()


def answers(sa, sb, minl=20):
    idx = ref.index_from_samples([sa, sb], bits=32)
    mums = np.asarray([(l, ab[0], ab[1]) for l, ab, _ in idx.getmums(minl)], dtype=np.int32).reshape(-1, 3)
    return dict(n=np.int64(idx.n), sa=digest(idx.SA), sai=digest(idx.SAi), lcp=digest(idx.LCP), mums=mums)


def main():
    A, B = fasta("123a.fa"), fasta("123b.fa")
    assert len(A) == 3 and len(B) == 3
    out = {}
    for name, seqs in (("a", A), ("b", B)):
        for k, s in enumerate(seqs):
            p, ln, ep, ec = pack(s)
            out["%s%d_packed" % (name, k)] = p
            out["%s%d_len" % (name, k)] = np.int64(ln)
            out["%s%d_exc_pos" % (name, k)] = ep
            out["%s%d_exc_chr" % (name, k)] = ec
    for tag, sa, sb in (("1", A[:1], B[:1]), ("2", A[1:2], B[1:2]), ("3", A[2:], B[2:]), ("123", A, B)):
        r = answers(sa, sb)
        print(tag, "n", int(r["n"]), "mums@20", len(r["mums"]))
        for k, v in r.items():
            out["ans%s_%s" % (tag, k)] = v
    os.makedirs(os.path.join(HERE, "real"), exist_ok=True)
    path = os.path.join(HERE, "real", "asp_niger.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
