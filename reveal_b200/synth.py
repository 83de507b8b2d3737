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


# ---- repeat-bearing inputs ---------------------------------------------------------------------------------
def repeat_ancestor(length, seed=1, families=6, tandems=40, segdups=3):
    """Ancestor codes (0..3) with the repeat structure of a real genome, deterministic in `seed`:
    interspersed families (50-500 copies of a 300-6000 bp element, copies 0-5 % diverged, some identical),
    tandem arrays (2-60 bp unit x 10-2000), a few segmental duplications of 10-50 kbp (0-1 % diverged).
    Roughly 10-15 % of the sequence ends up repetitive."""
    rng = np.random.default_rng(seed * 7919 + 13)
    g = rng.integers(0, 4, size=length, dtype=np.uint8)

    def put(at, piece):
        at = int(at)
        piece = piece[: max(0, length - at)]
        g[at:at + len(piece)] = piece

    def noisy(piece, div):
        if div <= 0:
            return piece
        p = piece.copy()
        m = rng.random(len(p)) < div
        p[m] = (p[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
        return p

    budget = length // 10
    for f in range(families):
        unit = rng.integers(0, 4, size=int(rng.integers(300, 6001)), dtype=np.uint8)
        copies = int(min(rng.integers(50, 501), max(2, budget // families // len(unit))))
        for c in range(copies):
            div = 0.0 if c % 5 == 0 else float(rng.random()) * 0.05   # every fifth copy is identical to the element
            put(rng.integers(0, length), noisy(unit, div))
    for t in range(tandems):
        unit = rng.integers(0, 4, size=int(rng.integers(2, 61)), dtype=np.uint8)
        reps = int(rng.integers(10, 2001))
        arr = np.tile(unit, reps)[: max(100, length // 200)]
        put(rng.integers(0, length), arr)
    for s in range(segdups):
        ln = int(min(rng.integers(10000, 50001), length // 8))
        src = int(rng.integers(0, max(1, length - ln)))
        put(rng.integers(0, length), noisy(g[src:src + ln].copy(), 0.01 * (s % 2)))
    return g


def repeat_genomes(n_genomes, length, seed=1, snp=0.01, indel=0.001, n_runs=4):
    """Like genomes(), on a repeat-bearing ancestor, with `n_runs` runs of 100-5000 'N' per genome."""
    g0 = repeat_ancestor(length, seed)
    out = []
    for k in range(n_genomes):
        rng = np.random.default_rng(seed * 1000 + 100 + k)
        codes = g0 if k == 0 else mutate(g0, rng, snp, indel)
        s = _ACGT[codes].copy()
        for _ in range(n_runs):
            ln = int(rng.integers(100, 5001))
            at = int(rng.integers(0, max(1, len(s) - ln)))
            s[at:at + ln] = ord("N")
        out.append(s)
    return out


def repeat_workload(n_genomes, length, seed=1):
    gs = repeat_genomes(n_genomes, length, seed)
    T, nsep = concat([[g] for g in gs])
    return T, nsep, n_genomes


def graph_like_workload(n_segments=500_000, mean_len=20, seed=1):
    """Two samples cut into `n_segments` short contigs each (a '$' after every one): the text shape a graph
    input gives the index (one sentinel per node)."""
    total = n_segments * mean_len
    a, b = genomes(2, total, seed)
    rng = np.random.default_rng(seed + 77)
    out = []
    for s in (a, b):
        cuts = np.unique(rng.integers(1, len(s), size=n_segments - 1))
        out.append(np.split(s, cuts))
    T, nsep = concat(out)
    return T, nsep, 2


def load_packed_fixture(path, pairs=(0, 1, 2)):
    """The real-data fixture written by tests/golden/make_real_golden.py: returns ([contigs of a], [contigs of b]) as
    uint8 ASCII arrays for the chosen contig numbers, and the npz handle (reference answers under 'ans*')."""
    z = np.load(path)
    out = []
    for name in ("a", "b"):
        seqs = []
        for k in pairs:
            p = z["%s%d_packed" % (name, k)]
            ln = int(z["%s%d_len" % (name, k)])
            c = np.empty((len(p), 4), np.uint8)
            for j in range(4):
                c[:, j] = (p >> (2 * j)) & 3
            s = _ACGT[c.reshape(-1)[:ln]].copy()
            s[z["%s%d_exc_pos" % (name, k)]] = z["%s%d_exc_chr" % (name, k)]
            seqs.append(s)
        out.append(seqs)
    return out[0], out[1], z
