"""A fixed number of cases of every leg of scripts/gpu_fuzz.py (seeded; the time-bounded runs of the same loop on the B200 box
are summarised under profiles/): index arrays + all sweeps against the oracle, align() against the reference's C aligner,
alignment graphs of the REM driver on this library against the same driver on the reference's extension."""
import importlib.util
import os

import numpy as np
import pytest

import oracle.ref as R

HERE = os.path.dirname(os.path.abspath(__file__))


def _fuzz():
    spec = importlib.util.spec_from_file_location("gpu_fuzz", os.path.join(os.path.dirname(HERE), "scripts", "gpu_fuzz.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.gpu
@pytest.mark.parametrize("leg,cases", [("tiny", 60), ("mid", 12)])
def test_cuda_fuzz_index_and_sweeps(cuda_lib, leg, cases):
    Z = _fuzz()
    done, trial = 0, 0
    while done < cases:
        rng = np.random.default_rng(77000 + trial)
        trial += 1
        r = Z.leg_tiny(cuda_lib, rng) if leg == "tiny" else Z.leg_mid(cuda_lib, rng)
        done += r is not None


@pytest.mark.gpu
@pytest.mark.skipif(not R.available(), reason="oracle/_ref (compiled reference) not present")
def test_cuda_fuzz_align_and_rem(tmp_path):
    from reveal_b200 import reveallib
    Z = _fuzz()
    for trial in range(6):
        Z.leg_align(reveallib, np.random.default_rng(78000 + trial))
    for trial in range(4):
        Z.leg_rem(reveallib, np.random.default_rng(79000 + trial), str(tmp_path), trial)
