"""Loader for the UNMODIFIED reference extension built into oracle/_ref/
(TEST INFRASTRUCTURE ONLY; see oracle/ref/Makefile).

`/root/reference` exists only in the build container; on the GPU box the
prebuilt oracle/_ref/*.so travel with the repo snapshot and are just loaded.
"""
import ctypes
import importlib.util
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(_HERE, "_ref")
REFERENCE_ROOT = os.environ.get("REVEAL_REFERENCE_ROOT", "/root/reference")
_mods = {}


def build():
    """(Re)build oracle/_ref from the reference sources when they are present."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "reveallib")):
        return False
    subprocess.check_call(["make", "-s", "-C", os.path.join(_HERE, "ref"), "REF=" + REFERENCE_ROOT])
    return True


def available(bits=32):
    return os.path.exists(os.path.join(_OUT, "_reveallib%s_ref.so" % ("64" if bits == 64 else "")))


def module(bits=32):
    """The reference's own `reveallib` (bits=32) / `reveallib64` (bits=64) module, or None."""
    name = "_reveallib%s_ref" % ("64" if bits == 64 else "")
    if name not in _mods:
        path = os.path.join(_OUT, name + ".so")
        if not os.path.exists(path):
            return None
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _mods[name] = mod
    return _mods[name]


def index_from_samples(samples, bits=32, rc=0, construct=True):
    """Feed `samples` (list of lists of sequences) through the reference's own
    addsample/addsequence/construct (interface.c:18-95,160-291)."""
    m = module(bits)
    idx = m.index()
    for k, seqs in enumerate(samples):
        idx.addsample("s%d" % k)
        for s in seqs:
            idx.addsequence(s if isinstance(s, str) else bytes(s).decode("ascii"))
    if construct:
        idx.construct(rc=rc) if rc else idx.construct()
    return idx


# --- raw entry points of the reference objects, for timing without the py-list getters ---

def _cdll(bits=32):
    name = "_reveallib%s_ref" % ("64" if bits == 64 else "")
    module(bits)  # make sure libpython symbols are bound through a normal import first
    return ctypes.CDLL(os.path.join(_OUT, name + ".so"))


def raw_build(T, bits=32):
    """divsufsort + ISA fill + compute_lcp exactly as construct() chains them
    (interface.c:213-253) but on numpy buffers: returns (SA, SAi, LCP) and the
    three phase times in seconds.  Used for the cpu_baseline timing."""
    import time
    L = _cdll(bits)
    T = np.ascontiguousarray(T, dtype=np.uint8)
    n = len(T)
    Tz = np.concatenate([T, np.zeros(1, dtype=np.uint8)])  # NUL terminator (interface.c:84)
    it = np.int64 if bits == 64 else np.int32
    lt = np.uint32 if bits == 64 else np.int32
    SA = np.empty(n, dtype=it)
    SAi = np.empty(n, dtype=it)
    LCP = np.empty(n, dtype=lt)
    fn = L.divsufsort64 if bits == 64 else L.divsufsort
    fn.restype = ctypes.c_int
    nn = ctypes.c_int64(n) if bits == 64 else ctypes.c_int32(n)
    t0 = time.perf_counter()
    rcode = fn(Tz.ctypes.data_as(ctypes.c_void_p), SA.ctypes.data_as(ctypes.c_void_p), nn)
    t1 = time.perf_counter()
    if rcode != 0:
        raise RuntimeError("divsufsort failed")
    SAi[SA] = np.arange(n, dtype=it)
    t2 = time.perf_counter()
    L.compute_lcp.restype = ctypes.c_int
    L.compute_lcp(Tz.ctypes.data_as(ctypes.c_void_p), SA.ctypes.data_as(ctypes.c_void_p),
                  SAi.ctypes.data_as(ctypes.c_void_p), LCP.ctypes.data_as(ctypes.c_void_p), nn)
    t3 = time.perf_counter()
    return SA, SAi, LCP, (t1 - t0, t2 - t1, t3 - t2)
