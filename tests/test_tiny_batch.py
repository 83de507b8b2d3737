"""reveallib.getmums_batch / rv_mums_tiny_batch (csrc/rv_tiny.cu): the flank pairs `extend` of finish / transform indexes one by one
(reveal/transformold.py:1170-1240), all in one launch -- every pair's list equals index.getmums of that pair on its own
(reference extension when present, else the oracle)."""
import numpy as np
import pytest

import oracle.port as P
import oracle.ref as R


def flank_pairs(rng, count):
    al = np.frombuffer(b"ACGT", np.uint8)
    pairs = []
    for k in range(count):
        la, lb = int(rng.integers(1, 201)), int(rng.integers(1, 201))
        a = al[rng.integers(0, 4, size=la)]
        if k % 3 == 0:      # homologous flanks: long shared stretches
            b = a.copy()[:lb] if lb <= la else np.concatenate([a, al[rng.integers(0, 4, size=lb - la)]])
            m = rng.random(len(b)) < 0.04
            b[m] = al[rng.integers(0, 4, size=int(m.sum()))]
        else:
            b = al[rng.integers(0, 4, size=lb)]
        a, b = a.tobytes().decode(), b.tobytes().decode()
        if k % 7 == 0:
            a = a[: len(a) // 2] + "N" * min(3, len(a)) + a[len(a) // 2:]
        pairs.append((a, b))
    pairs.append(("ACGT" * 50, "ACGT" * 50))            # tandem: no unique match
    pairs.append(("A", "A"))
    pairs.append(("ACGTTGCAAGGCTTAACCGGTTAAC" * 30, "ACGTTGCAAGGCTTAACCGGTTAAC" * 30))   # longer than the block path: regular build
    return pairs


def expected(pairs, minl):
    out = []
    for a, b in pairs:
        if R.available():
            idx = R.index_from_samples([[a], [b]])
            out.append([tuple(m) for m in idx.getmums(minl)])
        else:
            T, nsep, _ = P.assemble([[a.encode()], [b.encode()]])
            out.append([(int(l), (int(x), int(y)), 0) for l, x, y in P.Index(T, nsep, 2).getmums(minl).tolist()])
    return out


@pytest.mark.parametrize("minl", [8, 20])
def test_getmums_batch_emulated(emu_reveallib, minl):
    pairs = flank_pairs(np.random.default_rng(minl), 40)
    got = emu_reveallib.mod32.getmums_batch(pairs, minl)
    assert got == expected(pairs, minl)
    assert sum(len(g) for g in got) > 10


@pytest.mark.gpu
def test_getmums_batch_cuda():
    from reveal_b200 import reveallib
    pairs = flank_pairs(np.random.default_rng(77), 3000)
    got = reveallib.getmums_batch(pairs, 20)
    assert got == expected(pairs, 20)
    assert reveallib.getmums_batch([], 20) == []
