"""Builds reveal_b200/libreveal_b200.so from reveal_b200/csrc/*.cu with nvcc for sm_100a.

In-tree on purpose: the .so travels with the repository snapshot to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
OUT = os.path.join(_HERE, "libreveal_b200.so")
OBJ = os.path.join(_HERE, "csrc", "build")
UNITS = ["rv_api", "rv_sa", "rv_lcp", "rv_sweep", "rv_split", "rv_chain", "rv_tiny"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
EXTRA = os.environ.get("RV_NVCC_EXTRA", "").split()
NVCC_FLAGS = EXTRA + ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _newest_source():
    t = 0.0
    for root in (CSRC, os.path.join(_HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force=False, verbose=False):
    """Compile every translation unit and link the shared library. Returns its path."""
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest_source():
        return OUT
    os.makedirs(OBJ, exist_ok=True)

    def cc(u):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, u + ".cu"), "-o", os.path.join(OBJ, u + ".o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (u, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(len(UNITS)) as ex:
        logs = list(ex.map(cc, UNITS))
    if verbose:
        sys.stderr.write("\n".join(logs))
    link = [NVCC, "-shared", "-o", OUT] + [os.path.join(OBJ, u + ".o") for u in UNITS] + ["-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return OUT


def build_extension(force=False):
    """Compile the CPython extension modules reveal_b200/reveallib*.so and reveallib64*.so (host side of the drop-in)."""
    import sysconfig
    src = os.path.join(CSRC, "ext", "reveallib_module.cpp")
    inc = sysconfig.get_paths()["include"]
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    outs = []
    for name, defs in (("reveallib", []), ("reveallib64", ["-DSA64=1"])):
        out = os.path.join(_HERE, name + suffix)
        outs.append(out)
        deps = (src, os.path.join(CSRC, "ext", "chain_dp.h"), os.path.join(_HERE, "..", "include", "reveal_b200.h"))
        if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
            continue
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-I" + inc] + defs + [src, "-o", out, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for %s:\n%s\n%s" % (name, r.stdout, r.stderr))
    # the alignment graph of the REM driver (host code, no CUDA)
    src = os.path.join(CSRC, "ext", "remcore_module.cpp")
    out = os.path.join(_HERE, "remcore" + suffix)
    outs.append(out)
    if force or not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(CSRC, "ext", "chain_dp.h"))):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-I" + inc, src, "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed for remcore:\n%s\n%s" % (r.stdout, r.stderr))
    return outs


if __name__ == "__main__":
    print(build_extension(force="--force" in sys.argv))
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
