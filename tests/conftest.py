import glob
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


@pytest.fixture(scope="session")
def emu_lib():
    """The product's CUDA sources compiled for the host with the fiber emulation
    (tests/emu) -- a logic checker for the kernels, usable without a GPU.
    TEST INFRASTRUCTURE: never reachable from the product loader."""
    from reveal_b200 import _native
    out = subprocess.run([os.path.join(ROOT, "tests", "emu", "build_emu.sh")], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.fail("emu build failed:\n" + out.stdout + out.stderr)
    return _native.bind(os.path.join(ROOT, "tests", "emu", "_build", "libreveal_emu.so"))


@pytest.fixture()
def emu_reveallib(emu_lib, monkeypatch):
    """reveal_b200.reveallib with the emulated kernels injected in place of the CUDA library."""
    from reveal_b200 import _native, reveallib
    monkeypatch.setattr(_native, "_lib", emu_lib)
    return reveallib


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a real GPU (gpu-marked tests only)."""
    from reveal_b200 import _native
    return _native.lib()
