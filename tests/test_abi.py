"""The C-ABI shared library loads and exports every symbol include/reveal_b200.h
declares (no compute calls: there is no GPU in the CPU test tier)."""
import os
import re

import pytest

from reveal_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "reveal_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rv_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_native.SIGNATURES)


def test_library_builds_loads_and_exports_everything():
    path = build.build()
    L = _native.bind(path)  # raises AttributeError on a missing symbol
    for name in declared_symbols():
        assert hasattr(L, name)
    assert b"sm_100a" in L.rv_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes
    L = _native.bind(build.build())
    h = ctypes.c_void_p()
    assert L.rv_index_create(ctypes.byref(h), None) != 0
    assert L.rv_last_error()


def test_product_loader_never_points_at_the_emulation():
    assert _native.LIB_PATH.endswith(os.path.join("reveal_b200", "libreveal_b200.so"))
    src = open(os.path.join(ROOT, "reveal_b200", "_native.py")).read()
    assert "emu" not in src and "oracle" not in src


def test_extension_modules_build_and_import():
    """The compiled drop-in modules (csrc/ext/reveallib_module.cpp) build, import and expose the reference's surface."""
    outs = build.build_extension()
    assert len(outs) == 3 and all(os.path.exists(o) for o in outs)   # reveallib, reveallib64, remcore
    from reveal_b200 import remcore, reveallib, reveallib64
    assert callable(reveallib.chain_dp) and all(hasattr(remcore.Graph, m) for m in ("graphalign", "pick", "coords", "export"))
    for mod in (reveallib, reveallib64):
        for name in ("addsample", "addsequence", "construct", "align", "splitindex", "getmums", "getmultimums", "getmultimems", "copy",
                     "n", "depth", "nsamples", "samples", "nodes", "leftnode", "rightnode", "nsep", "SA", "SAi", "SO", "LCP", "T"):
            assert hasattr(mod.index, name), name
        assert issubclass(mod.error, Exception)
        idx = mod.index(sa="", lcp="", cache=0)
        idx.addsample("a")
        assert idx.addsequence("ACGT") == (0, 4) and idx.n == 5 and idx.nsamples == 1
