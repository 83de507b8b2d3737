#!/usr/bin/env python
"""Builds tuning variants of the CUDA library into build/variants/<name>.so (git-ignored, shipped by gpurun):
    build_variants.py name1="-DRV_PR_MINBLOCKS=8" name2="-DRV_PR_CHUNK=128 -DRV_PR_MINBLOCKS=12" ..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from reveal_b200 import build as B  # noqa: E402

OUT = os.path.join(ROOT, "build", "variants")


def one(spec):
    name, flags = spec.split("=", 1)
    d = os.path.join(OUT, name + "_obj")
    os.makedirs(d, exist_ok=True)

    def cc(u):
        cmd = [B.NVCC] + B.NVCC_FLAGS + flags.split() + ["-c", os.path.join(B.CSRC, u + ".cu"), "-o", os.path.join(d, u + ".o")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
    with ThreadPoolExecutor(len(B.UNITS)) as ex:
        list(ex.map(cc, B.UNITS))
    so = os.path.join(OUT, name + ".so")
    subprocess.check_call([B.NVCC, "-shared", "-o", so] + [os.path.join(d, u + ".o") for u in B.UNITS] + ["-cudart", "static"])
    return so


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        print(one(spec))
