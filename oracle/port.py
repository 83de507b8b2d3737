"""ctypes wrapper over oracle/reveal_oracle.c (CPU oracle, TEST INFRASTRUCTURE ONLY).

Function names follow the reference symbols they restate
(reveallib/interface.c, reveallib/reveal.c); see the C file for file:line.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "reveal_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u16p = ctypes.POINTER(ctypes.c_uint16)


def build(force=False):
    """Compile the C restatement (gcc -O2, like distutils would for the reference)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB, _SRC])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.orc_suffix_array.restype = ctypes.c_int
        L.orc_suffix_array.argtypes = [_u8p, ctypes.c_int64, _i32p]
        L.orc_inverse.restype = None
        L.orc_inverse.argtypes = [_i32p, ctypes.c_int64, _i32p]
        L.orc_compute_lcp.restype = None
        L.orc_compute_lcp.argtypes = [_u8p, _i32p, _i32p, _i32p, ctypes.c_int64]
        L.orc_build_so.restype = None
        L.orc_build_so.argtypes = [_i64p, ctypes.c_int32, ctypes.c_int64, _u16p]
        L.orc_comp_table.restype = None
        L.orc_comp_table.argtypes = [_u8p]
        L.orc_revcomp.restype = None
        L.orc_revcomp.argtypes = [_u8p, ctypes.c_int64]
        L.orc_getmums.restype = ctypes.c_int64
        L.orc_getmums.argtypes = [_u8p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                  ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _i64p, ctypes.c_int64]
        L.orc_getmulti.restype = ctypes.c_int
        L.orc_getmulti.argtypes = [_u8p, _i32p, _i32p, _u16p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32,
                                   ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _i64p, ctypes.c_int64,
                                   _i64p, ctypes.c_int64, _i64p, _i64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def as_text(T):
    """bytes / str / uint8 array -> contiguous uint8 array (a private copy)."""
    if isinstance(T, str):
        T = T.encode("ascii")
    if isinstance(T, (bytes, bytearray)):
        return np.frombuffer(bytes(T), dtype=np.uint8).copy()
    return np.ascontiguousarray(T, dtype=np.uint8).copy()


def suffix_array(T):
    T = as_text(T)
    SA = np.empty(len(T), dtype=np.int32)
    if lib().orc_suffix_array(_p(T, _u8p), len(T), _p(SA, _i32p)) != 0:
        raise MemoryError("oracle suffix_array failed")
    return SA


def inverse(SA):
    SA = np.ascontiguousarray(SA, dtype=np.int32)
    SAi = np.empty_like(SA)
    lib().orc_inverse(_p(SA, _i32p), len(SA), _p(SAi, _i32p))
    return SAi


def compute_lcp(T, SA, SAi):
    T = as_text(T)
    SA = np.ascontiguousarray(SA, dtype=np.int32)
    SAi = np.ascontiguousarray(SAi, dtype=np.int32)
    LCP = np.empty(len(T), dtype=np.int32)
    lib().orc_compute_lcp(_p(T, _u8p), _p(SA, _i32p), _p(SAi, _i32p), _p(LCP, _i32p), len(T))
    return LCP


def build_so(nsep, nsamples, n):
    nsep = np.ascontiguousarray(nsep, dtype=np.int64)
    SO = np.empty(n, dtype=np.uint16)
    lib().orc_build_so(_p(nsep, _i64p), nsamples, n, _p(SO, _u16p))
    return SO


def comp_table():
    t = np.zeros(128, dtype=np.uint8)
    lib().orc_comp_table(_p(t, _u8p))
    return t


def revcomp_inplace(T, start):
    """Reverse-complement T[start:] in place (interface.c:168-172)."""
    sub = np.ascontiguousarray(T[start:])
    lib().orc_revcomp(_p(sub, _u8p), len(sub))
    T[start:] = sub


def getmums(T, SA, LCP, nsep0, minl, rc=0, nT=None, rem=False):
    """reveal.c:55-116 (rem=False) / :119-180 (rem=True). Returns int64 [k,3] rows (l,a,b)."""
    T = as_text(T)
    SA = np.ascontiguousarray(SA, dtype=np.int32)
    LCP = np.ascontiguousarray(LCP, dtype=np.int32)
    n = len(SA)
    nT = len(T) if nT is None else nT
    cap = 1024
    while True:
        out = np.empty((cap, 3), dtype=np.int64)
        k = lib().orc_getmums(_p(T, _u8p), _p(SA, _i32p), _p(LCP, _i32p), n, nT, int(nsep0), int(rc),
                              int(minl), int(bool(rem)), _p(out, _i64p), cap)
        if k <= cap:
            return out[:k].copy()
        cap = int(k)


def getmulti(T, SA, LCP, SO, nsep0, main_nsamples, minl=0, minn=2, mem=False):
    """reveal.c:436-580 getmultimums (mem=False) / :292-434 getmultimems (mem=True).

    Returns (hdr int64 [k,3] rows (l, count_field, first_member), members int64 [m,2] rows (sample,pos))."""
    T = as_text(T)
    SA = np.ascontiguousarray(SA, dtype=np.int32)
    LCP = np.ascontiguousarray(LCP, dtype=np.int32)
    so_p = None
    if SO is not None:
        SO = np.ascontiguousarray(SO, dtype=np.uint16)
        so_p = _p(SO, _u16p)
    hc, mc = 1024, 4096
    nr, nm = ctypes.c_int64(0), ctypes.c_int64(0)
    while True:
        hdr = np.empty((hc, 3), dtype=np.int64)
        mem_ = np.empty((mc, 2), dtype=np.int64)
        rcode = lib().orc_getmulti(_p(T, _u8p), _p(SA, _i32p), _p(LCP, _i32p), so_p, len(SA), int(nsep0),
                                   int(main_nsamples), int(minl), int(minn), int(bool(mem)), _p(hdr, _i64p), hc,
                                   _p(mem_, _i64p), mc, ctypes.byref(nr), ctypes.byref(nm))
        if rcode != 0:
            raise MemoryError("oracle getmulti failed")
        if nr.value <= hc and nm.value <= mc:
            return hdr[:nr.value].copy(), mem_[:nm.value].copy()
        hc, mc = max(hc, nr.value), max(mc, nm.value)


def multi_to_tuples(hdr, members):
    """CSR -> the reference's Python shape [(l, n, ((sample,pos),...)), ...]."""
    out = []
    for k in range(len(hdr)):
        l, field, first = (int(x) for x in hdr[k])
        end = int(hdr[k + 1][2]) if k + 1 < len(hdr) else len(members)
        out.append((l, field, tuple((int(s), int(p)) for s, p in members[first:end])))
    return out


class Index:
    """Restatement of the `index` build pipeline, interface.c:160-291 (construct)."""

    def __init__(self, T, nsep, nsamples, rc=0):
        self.T = as_text(T)
        self.n = len(self.T)
        self.nT = self.n
        self.nsep = [int(x) for x in nsep]
        self.nsamples = int(nsamples)
        self.rc = int(rc)
        if self.rc == 1:
            revcomp_inplace(self.T, self.nsep[0])
        self.SA = suffix_array(self.T)
        self.SAi = inverse(self.SA)
        self.LCP = compute_lcp(self.T, self.SA, self.SAi)
        self.SO = build_so(self.nsep, self.nsamples, self.n) if self.nsamples > 2 else None

    def getmums(self, minl, rem=False):
        return getmums(self.T, self.SA, self.LCP, self.nsep[0], minl, rc=self.rc, nT=self.nT, rem=rem)

    def getmultimums(self, minlength=0, minn=2):
        return getmulti(self.T, self.SA, self.LCP, self.SO, self.nsep[0] if self.nsep else -1, self.nsamples,
                        minlength, minn, mem=False)

    def getmultimems(self, minlength=0, minn=2):
        return getmulti(self.T, self.SA, self.LCP, self.SO, self.nsep[0] if self.nsep else -1, self.nsamples,
                        minlength, minn, mem=True)


def assemble(samples):
    """Text assembly exactly like addsample/addsequence (interface.c:18-95).

    samples: list of samples, each a list of sequences (str/bytes).
    Returns (T uint8 array, nsep list, nodes list of (start,end))."""
    parts, nsep, nodes, n = [], [], [], 0
    for k, seqs in enumerate(samples):
        if k > 0:
            nsep.append(n - 1)
        for s in seqs:
            if isinstance(s, str):
                s = s.encode("ascii")
            parts.append(bytes(s) + b"$")
            nodes.append((n, n + len(s)))
            n += len(s) + 1
    return np.frombuffer(b"".join(parts), dtype=np.uint8).copy(), nsep, nodes
