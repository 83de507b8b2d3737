"""Shared helpers of the parity tests: run one build + sweeps through the C-ABI
(any library exporting it) and through the oracle, and compare bit-exactly."""
import ctypes
import os

import numpy as np

from reveal_b200 import _native


class NativeIndex:
    """Thin test-side wrapper over the C-ABI handle."""

    def __init__(self, L, T, nsep, nsamples, rc=0, device_ptr=None):
        self.L = L
        self.h = ctypes.c_void_p()
        _native.check(L, L.rv_index_create(ctypes.byref(self.h), None))
        self.T = np.ascontiguousarray(T, dtype=np.uint8)
        self.n = len(self.T)
        self.nsamples = int(nsamples)
        ns = np.asarray(nsep, dtype=np.int64)
        _native.check(L, L.rv_build(self.h, self.T.ctypes.data, self.n, ns.ctypes.data if len(ns) else None, self.nsamples, int(rc)))

    def rebuild(self, T, nsep, nsamples, rc=0):
        """Another build on the same handle (workspace and alphabet cache are reused)."""
        L = self.L
        self.T = np.ascontiguousarray(T, dtype=np.uint8)
        self.n = len(self.T)
        self.nsamples = int(nsamples)
        ns = np.asarray(nsep, dtype=np.int64)
        _native.check(L, L.rv_build(self.h, self.T.ctypes.data, self.n, ns.ctypes.data if len(ns) else None, self.nsamples, int(rc)))

    def close(self):
        if self.h is not None:
            self.L.rv_index_free(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def arr(self, which, bits=32):
        L = self.L
        if which == "SO":
            a = np.empty(self.n, np.uint16)
            _native.check(L, L.rv_get_so(self.h, a.ctypes.data))
            return a
        if which == "T":
            a = np.empty(self.n, np.uint8)
            _native.check(L, L.rv_get_text(self.h, a.ctypes.data))
            return a
        if which == "LCP":
            a = np.empty(self.n, np.uint32 if bits == 64 else np.int32)
            _native.check(L, L.rv_get_lcp(self.h, a.ctypes.data, bits))
            return a
        a = np.empty(self.n, np.int64 if bits == 64 else np.int32)
        _native.check(L, (L.rv_get_sa if which == "SA" else L.rv_get_sai)(self.h, a.ctypes.data, bits))
        return a

    def mums(self, minl, flavour=0):
        L = self.L
        c = ctypes.c_int64()
        _native.check(L, L.rv_mums_pair_count(self.h, int(minl), int(flavour), ctypes.byref(c)))
        rows = np.empty((c.value, 3), np.int64)
        _native.check(L, L.rv_mums_pair_fetch(self.h, rows.ctypes.data, c.value))
        return rows

    def multimums(self, minl, minn=2):
        L = self.L
        nr, nm = ctypes.c_int64(), ctypes.c_int64()
        _native.check(L, L.rv_mums_multi_count(self.h, int(minl), int(minn), ctypes.byref(nr), ctypes.byref(nm)))
        hdr = np.empty((nr.value, 3), np.int64)
        mem = np.empty((nm.value, 2), np.int64)
        _native.check(L, L.rv_mums_multi_fetch(self.h, hdr.ctypes.data, nr.value, mem.ctypes.data, nm.value))
        return hdr, mem

    def times(self):
        t = _native.Times()
        _native.check(self.L, self.L.rv_get_times(self.h, ctypes.byref(t)))
        return t.as_dict()


def assert_same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, "%s: shape %s vs %s" % (what, a.shape, b.shape)
    if a.size and not np.array_equal(a, b):
        bad = np.flatnonzero((a != b).reshape(len(a), -1).any(axis=1))
        raise AssertionError("%s: %d rows differ, first at %d: %s vs %s" % (what, len(bad), bad[0], a[bad[0]], b[bad[0]]))


def check_against_golden(L, g):
    """Build with library L on the golden's input and compare every array / MUM list bit-exactly."""
    ns = int(g["nsamples"])
    with NativeIndex(L, g["T_in"], g["nsep"], ns, rc=int(g["rc"])) as idx:
        assert_same(idx.arr("SA"), g["SA"], "SA")
        assert_same(idx.arr("SAi"), g["SAi"], "SAi")
        assert_same(idx.arr("LCP"), g["LCP"], "LCP")
        assert_same(idx.arr("T"), g["T_indexed"], "T")
        assert_same(idx.arr("SA", 64), g["SA"].astype(np.int64), "SA64")
        assert_same(idx.arr("LCP", 64), g["LCP"].astype(np.uint32), "LCP64")
        assert_same(idx.mums(int(g["minl"])), g["mums"], "getmums")
        if ns > 2:
            assert_same(idx.arr("SO"), g["SO"], "SO")
            hdr, mem = idx.multimums(int(g["minl"]), int(g["minn"]))
            assert_same(hdr, g["mm_hdr"], "getmultimums hdr")
            assert_same(mem, g["mm_mem"], "getmultimums members")


def check_against_oracle(L, T, nsep, nsamples, rc=0, minl=5, minn=2, arrays=True):
    """Build with library L and with the oracle port on the same input; compare bit-exactly."""
    import oracle.port as P
    o = P.Index(T, nsep, nsamples, rc)
    with NativeIndex(L, T, nsep, nsamples, rc=rc) as idx:
        if arrays:
            assert_same(idx.arr("SA"), o.SA, "SA")
            assert_same(idx.arr("SAi"), o.SAi, "SAi")
            assert_same(idx.arr("LCP"), o.LCP, "LCP")
            if nsamples > 2:
                assert_same(idx.arr("SO"), o.SO, "SO")
        if nsamples >= 2:
            for fl in (0, 1):
                assert_same(idx.mums(minl, fl), o.getmums(minl, rem=bool(fl)), "getmums flavour %d" % fl)
        hdr, mem = idx.multimums(minl, minn)
        oh, om = o.getmultimums(minl, minn)
        assert_same(hdr, oh, "getmultimums hdr")
        assert_same(mem, om, "getmultimums members")
        return idx.times(), len(oh)


def random_related(rng, nsamples, length, sigma=4, snp=0.02, alphabet=b"ACGT"):
    """nsamples noisy copies of one random sequence over the first `sigma` letters."""
    al = np.frombuffer(alphabet, np.uint8)
    base = rng.integers(0, sigma, size=length)
    out = []
    for _ in range(nsamples):
        g = base.copy()
        m = rng.random(length) < snp
        g[m] = rng.integers(0, sigma, size=int(m.sum()))
        out.append([al[g].tobytes()])
    return out


def check_handle_reuse(L):
    """Several builds on ONE handle: the second build starts speculatively with the first text's alphabet; a text
    with a symbol that alphabet lacks must be rebuilt, a text with fewer symbols must not be disturbed."""
    import oracle.port as P
    rng = np.random.default_rng(5)
    texts = [random_related(rng, 2, 3000, 4),                                # A C G T $
             random_related(rng, 2, 2500, 6, alphabet=b"ACGTNR"),             # + N R : superset -> redo
             random_related(rng, 3, 2000, 2, alphabet=b"AC"),                 # subset
             random_related(rng, 2, 2200, 16, alphabet=b"ACGTNRYKMSWBDHVn")]  # 17 symbols: byte comparison path
    idx = None
    for samples in texts:
        T, nsep, _ = P.assemble(samples)
        ns = len(samples)
        if idx is None:
            idx = NativeIndex(L, T, nsep, ns)
        else:
            idx.rebuild(T, nsep, ns)
        o = P.Index(T, nsep, ns)
        assert_same(idx.arr("SA"), o.SA, "SA")
        assert_same(idx.arr("SAi"), o.SAi, "SAi")
        assert_same(idx.arr("LCP"), o.LCP, "LCP")
        assert_same(idx.mums(5, 1), o.getmums(5, rem=True), "getmums")
    idx.close()


def check_sweep_prefetch(L):
    """Sweeps between rebuilds of one handle (result buffer, scratch and alphabet are reused; a build that enqueued the sweep
    its handle was asked for last in front of its own synchronisation point was measured and dropped -- it saved nothing -- but
    what such a short cut could get wrong stays tested): the answer must be the sweep of the NEW text whatever happens in
    between: another minimum length, a text that needs the doubling rounds, a changed text (rv_put_text), a request repeated."""
    import ctypes

    import oracle.port as P
    rng = np.random.default_rng(17)
    al = np.frombuffer(b"ACGT", np.uint8)
    rep = al[rng.integers(0, 4, size=40)].tobytes()
    base = al[rng.integers(0, 4, size=2500)].tobytes()
    texts = [random_related(rng, 2, 3000, 4, snp=0.02),
             random_related(rng, 2, 2600, 4, snp=0.03),
             [[base[:1200] + rep * 30 + base[1200:]], [base[:700] + b"T" + base[701:1900] + rep * 4 + base[1900:]]],   # stage 4
             random_related(rng, 2, 2000, 4, snp=0.02)]
    idx = None
    for k, samples in enumerate(texts):
        T, nsep, _ = P.assemble(samples)
        if idx is None:
            idx = NativeIndex(L, T, nsep, 2)
        else:
            idx.rebuild(T, nsep, 2)
        o = P.Index(T, nsep, 2)
        if k == 1:
            assert_same(idx.mums(9, 1), o.getmums(9, rem=True), "another minl right after a build")
        assert_same(idx.mums(6, 1), o.getmums(6, rem=True), "getmums text %d" % k)
        assert_same(idx.mums(6, 1), o.getmums(6, rem=True), "getmums text %d, asked again" % k)
        assert_same(idx.mums(6, 0), o.getmums(6), "other flavour")
        assert_same(idx.mums(6, 1), o.getmums(6, rem=True), "back to the hinted request")
    # a change of the text between the build and the request voids the prefetched answer
    T, nsep, _ = P.assemble(texts[0])
    idx.rebuild(T, nsep, 2)
    o = P.Index(T, nsep, 2)
    want = o.getmums(6, rem=True)
    pos = int(want[0][1]) - 1 if len(want) and want[0][1] > 0 else 5
    T2 = T.copy()
    T2[pos] = ord("a")   # lower case left of a match: left-maximal now whatever stood there
    _native.check(L, L.rv_put_text(idx.h, pos, T2[pos:pos + 1].ctypes.data, 1))
    o2 = P.Index(T, nsep, 2)
    o2.T = T2
    assert_same(idx.mums(6, 1), o2.getmums(6, rem=True), "after rv_put_text")
    idx.close()


def check_pack_block(L):
    """rv_result_pack_device into a peer block allocated by this process (rv_peer_alloc / rv_peer_read / rv_peer_free):
    header row (count, pack sequence number, 0) and the rows, with room to spare and with a capacity below the count."""
    import ctypes

    import oracle.port as P
    from reveal_b200 import _native
    rng = np.random.default_rng(21)
    T, nsep, _ = P.assemble(random_related(rng, 2, 4000, 4))
    idx = NativeIndex(L, T, nsep, 2)
    want = np.asarray(P.Index(T, nsep, 2).getmums(6, rem=True), dtype=np.int64).reshape(-1, 3)
    assert_same(idx.mums(6, 1), want, "getmums")
    k = len(want)
    assert k > 8
    p = ctypes.c_void_p()
    handle = (ctypes.c_uint8 * 64)()
    nbytes = (k + 9) * 24
    _native.check(L, L.rv_peer_alloc(nbytes, ctypes.byref(p), handle))
    assert p.value and any(handle)
    try:
        for seq, cap in ((1, k + 8), (2, 5), (3, 0)):
            _native.check(L, L.rv_result_pack_device(idx.h, p, cap))
            blk = np.full((k + 9) * 3, -7, dtype=np.int64)
            idx.arr("SA")  # a fetch synchronises the handle's stream
            _native.check(L, L.rv_peer_read(p, blk.ctypes.data, nbytes))
            m = min(k, cap)
            assert blk[:3].tolist() == [k, seq, 0]
            assert_same(blk[3:3 * (m + 1)].reshape(m, 3), want[:m], "packed rows")
    finally:
        _native.check(L, L.rv_peer_free(p))
    assert L.rv_peer_alloc(0, ctypes.byref(p), handle) != 0
    idx.close()


REAL_FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real", "asp_niger.npz")
REAL_PAIRS = {"1": (0,), "2": (1,), "3": (2,), "123": (0, 1, 2)}


def load_real(tag):
    """(T, nsep, answers) of one of the reference's real-data pairs (1a/1b, 2a/2b, 3a/3b, 123a/123b) from the committed
    compact fixture; answers = the unmodified reference's n, sha256 of SA / SAi / LCP and its getmums(20) rows."""
    from reveal_b200 import synth
    a, b, z = synth.load_packed_fixture(REAL_FIXTURE, REAL_PAIRS[tag])
    T, nsep = synth.concat([a, b])
    ans = {k: z["ans%s_%s" % (tag, k)] for k in ("n", "sa", "sai", "lcp", "mums")}
    return T, nsep, ans


def digest32(arr):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(np.asarray(arr, dtype=np.int32)).tobytes()).hexdigest()


def check_against_real(build, tag):
    """build(T, nsep) -> object with .SA/.SAi/.LCP arrays (callables or arrays) and mums(minl); compares with the reference's answers."""
    T, nsep, ans = load_real(tag)
    assert len(T) == int(ans["n"])
    SA, SAi, LCP, mums = build(T, nsep)
    assert digest32(SA) == str(ans["sa"]), "SA differs from the reference's"
    assert digest32(SAi) == str(ans["sai"]), "SAi differs from the reference's"
    assert digest32(LCP) == str(ans["lcp"]), "LCP differs from the reference's"
    assert_same(np.asarray(mums, dtype=np.int64).reshape(-1, 3), ans["mums"].astype(np.int64), "getmums(20)")
    return len(mums)
