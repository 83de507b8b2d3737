"""Recursion parity: reveal_b200's align() against the reference's unmodified
C aligner (oracle/_ref), both driven by the same deterministic callbacks.  Compared: every call the
mumpicker receives (depth, n, nsamples, nodes, the MUM list incl. order), every chosen MUM, and the
final text with its lower-cased matched regions.  CPU tier = emulated kernels on small inputs; the
gpu-marked cases run the CUDA library."""
import numpy as np
import pytest

import oracle.port as P
import oracle.ref as R
from align_callbacks import make_callbacks
from util import random_related

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref (compiled reference) not present")


def run_reference(samples, minl, minn, maxsteps=None):
    log = []
    idx = R.index_from_samples(samples)
    mp, ga = make_callbacks(log, minlen=minl, maxsteps=maxsteps)
    idx.align(mp, ga, threads=0, minl=minl, minn=minn)
    return log, idx.T[:idx.n]


def run_ours(reveallib, samples, minl, minn, maxsteps=None, threads=1):
    """threads=1: one step per launch in the reference's LIFO order; threads=0 (the default of the drop-in): frontier batching --
    the device parts of the steps waiting on the queue share one launch, so the sub-indexes are visited in another order."""
    log = []
    idx = reveallib.index()
    for k, seqs in enumerate(samples):
        idx.addsample("s%d" % k)
        for s in seqs:
            idx.addsequence(s if isinstance(s, str) else bytes(s).decode("ascii"))
    idx.construct()
    mp, ga = make_callbacks(log, minlen=minl, maxsteps=maxsteps)
    idx.align(mp, ga, threads=threads, minl=minl, minn=minn)
    return log, idx.T


def per_subindex(log):
    """The log as one record per visited sub-index: its pick event and, if a MUM was chosen, the align event that follows it."""
    out = []
    for e in log:
        if e[0] == "pick":
            out.append([e, None])
        else:
            out[-1][1] = e
    return sorted((repr(p), repr(a)) for p, a in out)


def compare(a, b, ordered=True):
    """ordered: the very same sequence of callback events; else the same SET of visited sub-indexes (every one with the same MUM
    list, choice and children) in any order -- what frontier batching promises, like the reference's own worker threads."""
    la, ta = a
    lb, tb = b
    assert len(la) == len(lb), "number of callback events differs: %d vs %d" % (len(la), len(lb))
    if ordered:
        for k, (x, y) in enumerate(zip(la, lb)):
            assert x == y, "event %d differs:\n ref  %s\n ours %s" % (k, str(x)[:600], str(y)[:600])
    else:
        for k, (x, y) in enumerate(zip(per_subindex(la), per_subindex(lb))):
            assert x == y, "sub-index record %d differs:\n ref  %s\n ours %s" % (k, str(x)[:600], str(y)[:600])
    assert ta == tb
    return len([e for e in la if e[0] == "align"])


CASES = [
    ("pair_small", 2, 1500, 4, 8, 2),
    ("pair_binary", 2, 800, 2, 10, 2),
    ("triple", 3, 1200, 4, 8, 2),
    ("five", 5, 700, 4, 7, 2),
    ("triple_minn3", 3, 1000, 4, 8, 3),
]


@needs_ref
@pytest.mark.parametrize("name,ns,length,sigma,minl,minn", CASES)
def test_align_matches_reference_emulated(emu_reveallib, name, ns, length, sigma, minl, minn):
    rng = np.random.default_rng(len(name) * 100 + length)
    samples = random_related(rng, ns, length, sigma, snp=0.03)
    ref = run_reference(samples, minl, minn)
    steps = compare(ref, run_ours(emu_reveallib, samples, minl, minn, threads=1))
    assert steps > 3
    assert compare(ref, run_ours(emu_reveallib, samples, minl, minn, threads=0), ordered=False) == steps   # batched frontier


@needs_ref
@pytest.mark.parametrize("small_maxn,bubble_maxn,bubble_cap", [("0", "100000", None), ("0", "0", None), ("600", "0", None),
                                                               ("16384", "100000", "32"), ("0", "100000", "32"), ("0", "0", "32")])
def test_align_general_step_paths_emulated(emu_reveallib, monkeypatch, small_maxn, bubble_maxn, bubble_cap):
    """RV_SMALL_MAXN / RV_BUBBLE_BLOCK_MAXN force the multi-kernel step and the grid-wide bubble detection; RV_BUBBLE_CAP=32
    makes matched intervals with more than 32 candidate slots take bubble_sort's windowed path (4096 in the product) in the
    single-block step, the one-block bubble kernel and the grid-wide detection + apply pair."""
    monkeypatch.setenv("RV_SMALL_MAXN", small_maxn)
    monkeypatch.setenv("RV_BUBBLE_BLOCK_MAXN", bubble_maxn)
    if bubble_cap:
        monkeypatch.setenv("RV_BUBBLE_CAP", bubble_cap)
    rng = np.random.default_rng(77)
    samples = random_related(rng, 3, 1100, 4, snp=0.03)
    assert compare(run_reference(samples, 8, 2), run_ours(emu_reveallib, samples, 8, 2)) > 3
    samples = random_related(rng, 2, 1500, 4, snp=0.03)
    assert compare(run_reference(samples, 8, 2), run_ours(emu_reveallib, samples, 8, 2)) > 3


def repeat_before_match(rng, rlen=300, mlen=400, nsamples=2):
    """Two samples whose longest MUM M starts right behind the second copy of a repeat R in sample 0, the first copy being
    followed by the first characters of M: every suffix inside that second copy matches its twin in the first copy ACROSS the
    start of M, i.e. bubble_sort of the leading child meets about rlen candidate slots for that matched interval."""
    al = np.frombuffer(b"ACGT", np.uint8)

    def rnd(k):
        return al[rng.integers(0, 4, size=k)].tobytes()
    R, M = rnd(rlen), rnd(mlen)
    s0 = rnd(500) + R + M[:25] + rnd(400) + R + M + rnd(300)
    return [[s0]] + [[rnd(700 - 60 * k) + M + rnd(350 + 40 * k)] for k in range(nsamples - 1)]


@needs_ref
@pytest.mark.parametrize("small_maxn,bubble_maxn", [("16384", "100000"), ("0", "100000"), ("0", "0")])
def test_bubble_sort_more_candidates_than_the_fast_path_holds_emulated(emu_reveallib, monkeypatch, small_maxn, bubble_maxn):
    """A matched interval with more candidate slots than bubble_sort's sorted fast path takes at once (RV_BUBBLE_CAP=32 here,
    4096 in the product) goes through the windowed replay -- in the single-block step, the one-block bubble kernel and the
    grid-wide detection + apply pair (round 1 handed this case to one thread).  The children's SA and LCP arrays are compared
    with the reference's own splitindex."""
    monkeypatch.setenv("RV_SMALL_MAXN", small_maxn)
    monkeypatch.setenv("RV_BUBBLE_BLOCK_MAXN", bubble_maxn)
    monkeypatch.setenv("RV_BUBBLE_CAP", "32")
    splitindex_case(emu_reveallib.mod32, 3, 0, 10, samples=repeat_before_match(np.random.default_rng(123), nsamples=3), min_steps=1)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("small_maxn,bubble_maxn", [("16384", "100000"), ("0", "100000"), ("0", "0")])
def test_bubble_sort_more_candidates_than_the_fast_path_holds_cuda(monkeypatch, small_maxn, bubble_maxn):
    """The same on the CUDA library at the product's limit: a 6000-character repeat in front of the match gives about 6000
    candidate slots, more than the 4096 the sorted fast path takes."""
    from reveal_b200 import reveallib
    monkeypatch.setenv("RV_SMALL_MAXN", small_maxn)
    monkeypatch.setenv("RV_BUBBLE_BLOCK_MAXN", bubble_maxn)
    splitindex_case(reveallib, 3, 0, 10, samples=repeat_before_match(np.random.default_rng(321), rlen=6000, mlen=700, nsamples=3), min_steps=1)


@needs_ref
def test_align_reference_test01_pair(emu_reveallib):
    """The reference's own in-memory pair (reveal/tests/test_reveal.py:36-41), minlength=1."""
    samples = [["ACTTGCTAGCTAGTCAG"], ["ACTAGCTAGCTAGTGAG"]]
    compare(run_reference(samples, 1, 2), run_ours(emu_reveallib, samples, 1, 2))


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("ns,length,minl,minn,maxsteps", [(2, 60000, 12, 2, None), (3, 40000, 12, 2, None), (5, 20000, 10, 2, None), (2, 400000, 20, 2, 400)])
def test_align_matches_reference_cuda(ns, length, minl, minn, maxsteps):
    from reveal_b200 import reveallib, synth
    gs = synth.genomes(ns, length, seed=9, snp=0.01, indel=0.001)
    samples = [[g.tobytes()] for g in gs]
    ref = run_reference(samples, minl, minn, maxsteps)
    steps = compare(ref, run_ours(reveallib, samples, minl, minn, maxsteps, threads=1))
    assert steps > 10
    if maxsteps is None:  # (a step budget makes the outcome depend on the visiting order)
        assert compare(ref, run_ours(reveallib, samples, minl, minn, None, threads=0), ordered=False) == steps   # batched frontier


def test_align_callback_failure_is_reported_and_leaves_no_wreckage(emu_reveallib):
    """A callback that raises in the middle of the recursion: align() raises (the reference sets err_flag and returns
    NULL with reveallib.error, interface.c:366-370,409-414), every queued child is released, and the module keeps
    working afterwards."""
    from align_callbacks import make_callbacks
    rng = np.random.default_rng(77)
    T, nsep, _ = P.assemble(random_related(rng, 2, 1200, 4))
    seqs = [s for s in bytes(T).decode().split("$") if s]

    def fresh():
        idx = emu_reveallib.index()
        for k, s in enumerate(seqs):
            idx.addsample("s%d" % k)
            idx.addsequence(s)
        idx.construct()
        return idx

    for bad in ("mumpicker", "graphalign", "garbage"):
        log = []
        mp, ga = make_callbacks(log, minlen=8)
        calls = {"n": 0}

        def picker(mums, idx, precomputed=False, minlength=0):
            calls["n"] += 1
            if bad == "mumpicker" and calls["n"] == 4:
                raise ValueError("picker broke")
            if bad == "garbage" and calls["n"] == 4:
                return 42  # not a tuple: "call to mumpicker failed" (reveal.c:839-847)
            return mp(mums, idx, precomputed=precomputed, minlength=minlength)

        def aligner(idx, mum):
            if bad == "graphalign" and calls["n"] >= 4:
                raise RuntimeError("graphalign broke")
            return ga(idx, mum)

        idx = fresh()
        with pytest.raises((emu_reveallib.error, ValueError, RuntimeError)):
            idx.align(picker, aligner, minl=8, minn=2)
        assert calls["n"] >= 4
        del idx
    # the module is still healthy: a complete alignment equals the reference's
    log_a, log_b = [], []
    idx = fresh()
    mp, ga = make_callbacks(log_a, minlen=8)
    idx.align(mp, ga, minl=8, minn=2)
    ref = R.index_from_samples([[s] for s in seqs])
    mp, ga = make_callbacks(log_b, minlen=8)
    ref.align(mp, ga, minl=8, minn=2)
    assert len(log_a) == len(log_b) and idx.T == ref.T[:ref.n]


@needs_ref
@pytest.mark.parametrize("ns,length,minl", [(3, 1000, 8), (4, 700, 7), (5, 500, 7)])   # the reference's getmultimums needs SO: N > 2
def test_splitindex_matches_reference(emu_reveallib, ns, length, minl):
    """index.splitindex (reveal.c:1515-1748): the recursion driven from Python, two levels deep, against the
    reference's own splitindex -- children's n / nsamples / depth / nodes / bounds / SA / LCP and the text."""
    splitindex_case(emu_reveallib.mod32, ns, length, minl)


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("ns,length,minl", [(3, 30000, 12), (5, 8000, 10), (4, 200000, 14)])
def test_splitindex_matches_reference_cuda(ns, length, minl):
    """The same on the CUDA library, three levels deep, with parents above and below the single-block limit."""
    from reveal_b200 import reveallib
    splitindex_case(reveallib, ns, length, minl)


def splitindex_case(mod32, ns, length, minl, samples=None, min_steps=3):
    rng = np.random.default_rng(900 + ns)
    if samples is None:
        samples = random_related(rng, ns, length, 4, snp=0.03)
    seqs = [[bytes(c).decode() for c in contigs] for contigs in samples]

    def build(mod):
        idx = mod.index()
        for k, contigs in enumerate(seqs):
            idx.addsample("s%d" % k)
            for c in contigs:
                idx.addsequence(c)
        idx.construct()
        return idx

    def describe(child):
        if child is None:
            return None
        return (child.n, child.nsamples, child.depth, sorted(child.nodes), child.leftnode, child.rightnode, list(child.SA), list(child.LCP))

    ours, ref = build(mod32), build(R.module(32))
    frontier = [(ours, ref)]
    steps = 0
    for level in range(3):
        nxt = []
        for a, b in frontier:
            ma = a.getmultimums(minlength=minl, minn=2)
            mb = b.getmultimums(minlength=minl, minn=2)
            assert [tuple(m) for m in ma] == [tuple(m) for m in mb]
            picks = []
            for idx, mums in ((a, ma), (b, mb)):
                mp, ga = make_callbacks([], minlen=minl)
                pick = mp(mums, idx)
                picks.append((pick, ga(idx, pick[0]) if pick else None))
            if not picks[0][0] or picks[0][1] is None:
                continue
            (pa, ra), (pb, rb) = picks
            assert ra == rb
            ka = a.splitindex(*ra, [], [])
            kb = b.splitindex(*rb, [], [])
            steps += 1
            for ca, cb in zip(ka, kb):
                assert describe(ca) == describe(cb)
                if ca is not None and ca.n > 1:
                    nxt.append((ca, cb))
        frontier = nxt
    assert steps >= min_steps
    assert ours.T == ref.T[:ref.n]


# ---- extract (reveal.c:1386-1505) -------------------------------------------------------------------------------------------
def extract_case(mod32, ns, length, minl, seed):
    """index.extract(intervals): the positions of a MUM (and of a second, shorter stretch) leave the index in place.  Against the
    reference's own extract: n, LCP, SAi, the text, and SA from slot 1 on -- the reference never writes SA[0] of its result
    (its copy loop starts at slot 1, reveal.c:1448); here slot 0 holds the suffix that belongs there, checked separately."""
    rng = np.random.default_rng(seed)
    samples = random_related(rng, ns, length, 4, snp=0.03)
    seqs = [[bytes(c).decode() for c in contigs] for contigs in samples]

    def build(mod):
        idx = mod.index()
        for k, contigs in enumerate(seqs):
            idx.addsample("s%d" % k)
            for c in contigs:
                idx.addsequence(c)
        idx.construct()
        return idx

    ours, ref = build(mod32), build(R.module(32))
    mums = ref.getmultimums(minlength=minl, minn=2) if ns > 2 else [(l, 2, ((0, a), (1, b))) for l, (a, b), _ in ref.getmums(minl)]
    best = max(mums, key=lambda m: m[0])
    l = best[0]
    intervals = [(p, p + l) for _, p in best[2]]
    sa_before = list(ours.SA)
    gone = set()
    for b, e in intervals:
        gone.update(range(b, e))
    ours.extract(list(intervals))
    ref.extract(list(intervals))
    assert ours.n == ref.n == len(sa_before) - len(gone)
    assert list(ours.LCP) == list(ref.LCP)
    sa_o, sa_r = list(ours.SA), list(ref.SA)
    assert sa_o[1:] == sa_r[1:]
    assert sa_o == [s for s in sa_before if s not in gone] or sa_o[0] == [s for s in sa_before if s not in gone][0]   # slot 0: the survivor that sorts first
    assert ours.T == ref.T[:len(ours.T)]
    keep = [p for p in range(len(sa_before)) if p not in gone]
    sai_o, sai_r = list(ours.SAi), list(ref.SAi)
    # (the reference lists the first n entries of its inverse array; positions >= n are not visible through the getter)
    # and never rewrites the inverse entry of the suffix that belongs into slot 0
    vis = [p for p in keep if p < ours.n and p != sa_o[0]]
    assert [sai_o[p] for p in vis] == [sai_r[p] for p in vis]
    if sa_o[0] < ours.n:
        assert sai_o[sa_o[0]] == 0
    # the sweeps of the extracted index run on its own arrays
    if ns == 2:
        o = [m for m in ours.getmums(minl)]
        assert all(not (set(range(a, a + ll)) & gone) and not (set(range(b, b + ll)) & gone) for ll, (a, b), _ in o)
    return len(intervals)


@needs_ref
@pytest.mark.parametrize("ns,length,minl", [(2, 1500, 8), (3, 1000, 8)])
def test_extract_matches_reference(emu_reveallib, ns, length, minl):
    assert extract_case(emu_reveallib.mod32, ns, length, minl, 40 + ns) >= 2


@needs_ref
@pytest.mark.gpu
@pytest.mark.parametrize("ns,length,minl", [(2, 40000, 14), (4, 15000, 12)])
def test_extract_matches_reference_cuda(ns, length, minl):
    from reveal_b200 import reveallib
    assert extract_case(reveallib, ns, length, minl, 60 + ns) >= 2


@needs_ref
@pytest.mark.parametrize("ns", [2, 3])
def test_align_mums_as_rows_emulated(emu_reveallib, ns):
    """align(mums_as_rows=True): pair MUM lists arrive as `mumrows` objects, multi-MUM lists (more than two samples) as
    `multimumrows` -- len(), iteration and indexing give the reference's tuples, so the same callbacks produce the same events;
    the buffer of a `mumrows` holds the int64 rows."""
    rng = np.random.default_rng(5)
    samples = random_related(rng, ns, 1500, 4, snp=0.03)
    ref = run_reference(samples, 8, 2)
    log = []
    idx = emu_reveallib.index()
    for k, seqs in enumerate(samples):
        idx.addsample("s%d" % k)
        for s in seqs:
            idx.addsequence(bytes(s).decode("ascii"))
    idx.construct()
    seen = []
    mp, ga = make_callbacks(log, minlen=8)

    def picker(mums, index, precomputed=False, minlength=0):
        seen.append(type(mums).__name__)
        if type(mums).__name__ == "mumrows" and len(mums):
            rows = np.frombuffer(memoryview(mums), dtype=np.int64).reshape(-1, 3)
            assert [(int(l), 2, ((0, int(a)), (1, int(b)))) for l, a, b in rows] == [tuple(m) for m in mums] == [mums[i] for i in range(len(mums))]
        return mp(mums, index, precomputed=precomputed, minlength=minlength)

    idx.align(picker, ga, threads=1, minl=8, minn=2, mums_as_rows=True)
    assert ("mumrows" if ns == 2 else "multimumrows") in seen
    assert compare(ref, (log, idx.T)) > 3


@needs_ref
def test_root_arrays_are_gone_after_align(emu_reveallib):
    """The reference frees the root's SA and LCP at its first split (reveal.c:1279-1284): a second align() stops with
    'Index not yet constructed' (interface.c:295-298) and the SA / LCP getters raise TypeError (interface.c:548-551), while
    SAi, T and n stay readable.  Same here -- a second align() would otherwise run on the inverse array the first one rewrote."""
    rng = np.random.default_rng(11)
    samples = random_related(rng, 2, 1200, 4, snp=0.03)

    def build(mod):
        idx = mod.index()
        for k, seqs in enumerate(samples):
            idx.addsample("s%d" % k)
            for s in seqs:
                idx.addsequence(bytes(s).decode("ascii"))
        idx.construct()
        return idx

    for mod, err in ((emu_reveallib, emu_reveallib.error), (R.module(32), R.module(32).error)):
        idx = build(mod)
        assert len(idx.SA) == idx.n
        mp, ga = make_callbacks([], minlen=8)
        idx.align(mp, ga, threads=0, minl=8, minn=2)
        with pytest.raises(err, match="not yet constructed"):
            idx.align(mp, ga, threads=0, minl=8, minn=2)
        with pytest.raises(TypeError, match="not yet constructed"):
            idx.SA
        with pytest.raises(TypeError, match="not yet constructed"):
            idx.LCP
        assert len(idx.SAi) == idx.n and len(idx.T) >= idx.n
    with pytest.raises(emu_reveallib.error):   # (the reference follows a NULL pointer here)
        idx2 = build(emu_reveallib)
        mp, ga = make_callbacks([], minlen=8)
        idx2.align(mp, ga, minl=8, minn=2)
        idx2.getmums(8)
    # an index whose root was never split keeps its arrays (no MUM of that length)
    idx3 = build(emu_reveallib)
    mp, ga = make_callbacks([], minlen=5000)
    idx3.align(mp, ga, minl=5000, minn=2)
    assert len(idx3.SA) == idx3.n
