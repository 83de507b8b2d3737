"""Deterministic synthetic genomes for parity tests and bench.py.

Recipe (mirrors the reference's simulator
utils/simulate.py:17-77 `mut`): ancestor = i.i.d. uniform ACGT from
default_rng(seed); descendant k>=1 = independent mutation of the ancestor with
default_rng(seed*1000+100+k): per-site SNP probability `snp` (uniform over the
three other bases), per-site indel probability `indel` (half insertions of
i.i.d. bases, half deletions; length ~ Zipf(1.7) capped at 2000).
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def ancestor(length, seed=1):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def mutate(codes, rng, snp=0.01, indel=0.001, maxindel=2000):
    """codes: uint8 array over {0,1,2,3}. Returns a mutated copy (uint8 codes)."""
    L = len(codes)
    u = rng.random(L)
    out = codes.copy()
    snp_at = np.nonzero(u < snp)[0]
    out[snp_at] = (out[snp_at] + rng.integers(1, 4, size=len(snp_at), dtype=np.uint8)) & 3
    ev = np.nonzero((u >= snp) & (u < snp + indel))[0]
    if len(ev) == 0:
        return out
    lens = np.minimum(rng.zipf(1.7, size=len(ev)), maxindel).astype(np.int64)
    is_ins = rng.random(len(ev)) < 0.5
    pieces, cur = [], 0
    for p, ln, ins in zip(ev.tolist(), lens.tolist(), is_ins.tolist()):
        if p < cur:
            continue  # swallowed by a previous deletion
        pieces.append(out[cur:p])
        if ins:
            pieces.append(rng.integers(0, 4, size=ln, dtype=np.uint8))
            cur = p
        else:
            cur = min(L, p + ln)
    pieces.append(out[cur:])
    return np.concatenate(pieces)


def genomes(n_genomes, length, seed=1, snp=0.01, indel=0.001):
    """List of `n_genomes` ASCII uint8 arrays (upper-case ACGT, no sentinel)."""
    g0 = ancestor(length, seed)
    out = [_ACGT[g0]]
    for k in range(1, n_genomes):
        rng = np.random.default_rng(seed * 1000 + 100 + k)
        out.append(_ACGT[mutate(g0, rng, snp, indel)])
    return out


def concat(seqs_per_sample):
    """Text assembly like index.addsample/addsequence (reference interface.c:18-95):
    every sequence is followed by '$'; nsep[k] = position of the last '$' of sample k.

    seqs_per_sample: list of samples, each a list of uint8 arrays / bytes.
    Returns (T uint8 array, nsep int64 array of len nsamples-1)."""
    parts, nsep, n = [], [], 0
    dollar = np.frombuffer(b"$", dtype=np.uint8)
    for k, seqs in enumerate(seqs_per_sample):
        if k > 0:
            nsep.append(n - 1)
        for s in seqs:
            a = np.frombuffer(bytes(s), dtype=np.uint8) if isinstance(s, (bytes, bytearray, str)) else np.asarray(s, dtype=np.uint8)
            parts.append(a)
            parts.append(dollar)
            n += len(a) + 1
    return np.concatenate(parts), np.asarray(nsep, dtype=np.int64)


def workload(n_genomes, length, seed=1, snp=0.01, indel=0.001):
    """(T, nsep, nsamples) for `n_genomes` single-contig genomes."""
    gs = genomes(n_genomes, length, seed, snp, indel)
    T, nsep = concat([[g] for g in gs])
    return T, nsep, n_genomes
