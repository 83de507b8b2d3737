"""`reveallib_ctypes` -- the ctypes twin of the compiled extension `reveal_b200.reveallib`
(csrc/ext/reveallib_module.cpp): the same mirror of the reference's CPython extension
(reveallib/interface.c + reveal.c) written in Python over the C-ABI, kept as the readable
binding and for environments without a C++ compiler.

Mirrors the `index` type of the reference (interface.c:841-881):

    methods  (interface.c:474-487)  addsample addsequence construct getmums
                                    getmultimums align copy ...
    getters  (interface.c:731-785)  n depth nsamples samples nodes leftnode
                                    rightnode nsep SA SAi SO LCP T
    ctor     (interface.c:515-518)  index(sa="", lcp="", cache=0)
    errors   (interface.c:933-936)  reveallib.error

Same names, argument meaning, return shapes and error behaviour, so code written
against the reference (rem.py, transform*.py, plot.py, chain.py) keeps working.
There is no CPU implementation behind this class: it fails loudly when the CUDA
library is missing or no GPU is present.
"""
import ctypes

import numpy as np

from . import _native

INT_MAX = 2 ** 31 - 1


class error(Exception):
    """reveallib.error (interface.c:933-936)."""


class index(object):
    _bits = 32  # reveallib: int32 saidx_t (reveal.h:11-12); reveallib64 overrides

    def __init__(self, sa="", lcp="", cache=0):
        # reveal_init, interface.c:489-521
        self._safile, self._lcpfile, self._cache = sa, lcp, int(cache)
        self._chunks = []          # pieces of T not yet concatenated
        self._Tarr = None          # uint8 array of the text (host copy)
        self._n = 0
        self._nT = 0
        self._nsep = []
        self._nsamples = 0
        self._rc = 0
        self._depth = 0
        self._built = False
        self._Tdirty = False       # device text differs from the host copy (align lower-cases matched bases)
        self._nmums = 0
        self._h = None
        self.samples = []
        self.nodes = set()
        self.leftnode = None
        self.rightnode = None
        self.skipmums = []
        self.main = None

    # ---- native handle ---------------------------------------------------------
    def _lib(self):
        return _native.lib()

    def _handle(self):
        if self._h is None:
            L = self._lib()
            h = ctypes.c_void_p()
            self._call(L.rv_index_create(ctypes.byref(h), None))
            self._h = h
        return self._h

    def _call(self, status):
        if status != 0:
            raise error(self._lib().rv_last_error().decode("utf-8", "replace"))

    def __del__(self):
        try:
            if self._h is not None:
                _native.lib().rv_index_free(self._h)
                self._h = None
        except Exception:
            pass

    # ---- text assembly (host logic; interface.c:18-95) ---------------------------
    def addsample(self, sample):
        if not isinstance(sample, str):
            raise error("Sample name has to be a string.")
        self.samples.append(sample)
        if self._nsamples > 0:
            self._nsep.append(self._n - 1)  # position of the last '$' of the previous sample (interface.c:42)
        self._nsamples += 1
        return None

    def addsequence(self, seq):
        if isinstance(seq, str):
            seq = seq.encode("ascii")
        elif not isinstance(seq, (bytes, bytearray)):
            raise TypeError("addsequence expects a str")
        l = len(seq)
        if self._bits == 32 and (self._n + (l + 1) + 1) > INT_MAX:  # interface.c:61-68
            raise error("Total amount of sequence too large, use \"reveal <subcommand> --64\" to use 64 bit suffix arrays instead.")
        s = self._n
        self._chunks.append(bytes(seq))
        self._chunks.append(b"$")
        self._n += l + 1
        self._Tarr = None
        intv = (s, self._n - 1)
        self.nodes.add(intv)
        return intv

    def _text(self):
        if self._Tarr is None:
            self._Tarr = np.frombuffer(b"".join(self._chunks), dtype=np.uint8).copy() if self._chunks else np.zeros(0, np.uint8)
            self._chunks = [self._Tarr.tobytes()] if self._n else []
        return self._Tarr

    # ---- construct (interface.c:160-291) -------------------------------------------
    def construct(self, rc=0):
        rc = 1 if rc == 1 else 0
        if rc and self._nsamples < 2:
            raise error("rc=1 needs a second sample.")
        if self._n == 0:
            raise error("No text to index.")  # interface.c:177-180
        L = self._lib()
        T = self._text()
        nsep = np.asarray(self._nsep, dtype=np.int64)
        h = self._handle()
        if self._safile:
            # precomputed arrays from the reference's cache files (interface.c:224-231, 255-262): raw saidx_t / lcp_t
            it = np.int64 if self._bits == 64 else np.int32
            sa = np.fromfile(self._safile, dtype=it, count=self._n)
            if len(sa) != self._n:
                raise error("suffix array file %s holds %d entries, expected %d" % (self._safile, len(sa), self._n))
            sa = np.ascontiguousarray(sa, dtype=np.int32)
            lcp = None
            if self._lcpfile:
                lcp = np.fromfile(self._lcpfile, dtype=np.uint32 if self._bits == 64 else np.int32, count=self._n)
                if len(lcp) != self._n:
                    raise error("lcp file %s holds %d entries, expected %d" % (self._lcpfile, len(lcp), self._n))
                lcp = np.ascontiguousarray(lcp, dtype=np.int32)
            self._call(L.rv_build_cached(h, T.ctypes.data, self._n, nsep.ctypes.data if len(nsep) else None, self._nsamples, rc,
                                         sa.ctypes.data, lcp.ctypes.data if lcp is not None else None))
        elif self._lcpfile:
            raise error("an lcp file needs its suffix array file (sa=...) as well")
        else:
            self._call(L.rv_build(h, T.ctypes.data, self._n, nsep.ctypes.data if len(nsep) else None, self._nsamples, rc))
        self._rc = rc
        self._nT = self._n
        if rc:
            # the reference reverse-complements its T in place (interface.c:168-172): mirror the device text
            self._call(L.rv_get_text(h, T.ctypes.data))
            self._chunks = [T.tobytes()]
        self._built = True
        if self._cache == 1:  # interface.c:182-189,273-285
            T.tofile(".reveal.t")
            self._array("SA").tofile(".reveal.sa")
            self._array("LCP").tofile(".reveal.lcp")
        self.main = self
        return None

    def times(self):
        """Device milliseconds of the last construct() per phase (not in the reference)."""
        t = _native.Times()
        self._call(self._lib().rv_get_times(self._handle(), ctypes.byref(t)))
        return t.as_dict()

    # ---- sweeps ----------------------------------------------------------------------
    def getmums_array(self, minl=0, flavour=0):
        """int64 [k,3] rows (l, a, b) in ascending SA rank (array form of getmums)."""
        if not self._built:
            raise error("Index not yet constructed.")
        L, h = self._lib(), self._handle()
        c = ctypes.c_int64()
        self._call(L.rv_mums_pair_count(h, int(minl), int(flavour), ctypes.byref(c)))
        rows = np.empty((c.value, 3), dtype=np.int64)
        self._call(L.rv_mums_pair_fetch(h, rows.ctypes.data, c.value))
        return rows

    def getmums(self, minl=0):
        """[(l, (a, b), rc), ...]  -- reveal.c:55-116."""
        rc = self._rc
        return [(l, (a, b), rc) for l, a, b in self.getmums_array(minl).tolist()]

    def getmultimums_arrays(self, minlength=0, minn=2):
        """(hdr int64 [k,3] rows (l, n, first_member), members int64 [m,2] rows (sample, pos))."""
        if not self._built:
            raise error("Index not yet constructed.")
        L, h = self._lib(), self._handle()
        nr, nm = ctypes.c_int64(), ctypes.c_int64()
        self._call(L.rv_mums_multi_count(h, int(minlength), int(minn), ctypes.byref(nr), ctypes.byref(nm)))
        hdr = np.empty((nr.value, 3), dtype=np.int64)
        mem = np.empty((nm.value, 2), dtype=np.int64)
        self._call(L.rv_mums_multi_fetch(h, hdr.ctypes.data, nr.value, mem.ctypes.data, nm.value))
        return hdr, mem

    def getmultimums(self, minlength=0, minn=2):
        """[(l, n, ((sample, pos), ...)), ...]  -- reveal.c:436-580."""
        hdr, mem = self.getmultimums_arrays(minlength, minn)
        mem = [tuple(x) for x in mem.tolist()]
        out = []
        for l, n, first in hdr.tolist():
            out.append((l, n, tuple(mem[first:first + n])))
        return out

    def getmultimems_arrays(self, minlength=0, minn=2):
        """(hdr int64 [k,3] rows (l, n_samples, first_member), members int64 [m,2]); a record's members run to the next
        record's first_member."""
        if not self._built:
            raise error("Index not yet constructed.")
        L, h = self._lib(), self._handle()
        nr, nm = ctypes.c_int64(), ctypes.c_int64()
        self._call(L.rv_mems_multi_count(h, int(minlength), int(minn), ctypes.byref(nr), ctypes.byref(nm)))
        hdr = np.empty((nr.value, 3), dtype=np.int64)
        mem = np.empty((nm.value, 2), dtype=np.int64)
        self._call(L.rv_mums_multi_fetch(h, hdr.ctypes.data, nr.value, mem.ctypes.data, nm.value))
        return hdr, mem

    def getmultimems(self, minlength=0, minn=2):
        """[(l, n_samples, ((sample, pos), ...)), ...]  -- reveal.c:292-434."""
        hdr, mem = self.getmultimems_arrays(minlength, minn)
        mem = [tuple(x) for x in mem.tolist()]
        out = []
        rows = hdr.tolist()
        for k, (l, c, first) in enumerate(rows):
            end = rows[k + 1][2] if k + 1 < len(rows) else len(mem)
            out.append((l, c, tuple(mem[first:end])))
        return out

    def copy(self):
        """A second index over the same text with its own SA / SAi / LCP / SO (interface.c:432-470).  The arrays are
        uploaded into the copy's device handle (no rebuild)."""
        if not self._built:
            raise error("Index not yet constructed.")
        new = type(self)()
        new.samples = list(self.samples)
        new.nodes = set(self.nodes)
        new._chunks = [self._text().tobytes()]
        new._n = self._n
        new._nsep = list(self._nsep)
        new._nsamples = self._nsamples
        T = new._text()
        nsep = np.asarray(new._nsep, dtype=np.int64)
        sa = np.ascontiguousarray(self._array("SA"), dtype=np.int32)
        lcp = np.ascontiguousarray(self._array("LCP"), dtype=np.int32)
        L = new._lib()
        # rc = 0: the text handed over is already the indexed (possibly reverse-complemented) one
        new._call(L.rv_build_cached(new._handle(), T.ctypes.data, new._n, nsep.ctypes.data if len(nsep) else None, new._nsamples, 0,
                                    sa.ctypes.data, lcp.ctypes.data))
        new._rc = self._rc
        new._nT = self._nT
        new._built = True
        new.main = new
        return new

    def align(self, mumpicker, align, threads=0, wpen=0, wscore=0, minl=0, minn=0):
        if not self._built:
            raise error("Index not yet constructed, alignment stopped.")  # interface.c:295-298
        from . import recursion
        return recursion.align(self, mumpicker, align, threads, wpen, wscore, minl, minn)

    # ---- getters (interface.c:539-785) --------------------------------------------------
    def _array(self, which):
        if not self._built:
            raise TypeError("Index not yet constructed.")
        L, h = self._lib(), self._handle()
        if which == "SO":
            if self._nsamples <= 2:
                raise TypeError("SO not available.")
            a = np.empty(self._n, dtype=np.uint16)
            self._call(L.rv_get_so(h, a.ctypes.data))
            return a
        if which == "LCP":
            a = np.empty(self._n, dtype=np.uint32 if self._bits == 64 else np.int32)
            self._call(L.rv_get_lcp(h, a.ctypes.data, self._bits))
            return a
        a = np.empty(self._n, dtype=np.int64 if self._bits == 64 else np.int32)
        self._call((L.rv_get_sa if which == "SA" else L.rv_get_sai)(h, a.ctypes.data, self._bits))
        return a

    SA = property(lambda self: self._array("SA").tolist())
    SAi = property(lambda self: self._array("SAi").tolist())
    LCP = property(lambda self: self._array("LCP").tolist())
    SO = property(lambda self: self._array("SO").tolist())
    n = property(lambda self: self._n)
    depth = property(lambda self: self._depth)
    nsamples = property(lambda self: self._nsamples)
    nsep = property(lambda self: list(self._nsep))

    @property
    def T(self):
        T = self._text()
        if self._built and self._Tdirty:  # matched regions are lower-cased in place during align (reveal.c:1230-1234)
            self._call(self._lib().rv_get_text(self._handle(), T.ctypes.data))
            self._chunks = [T.tobytes()]
            self._Tdirty = False
        return T.tobytes().decode("latin-1")


def _version():
    return _native.lib().rv_version().decode()
