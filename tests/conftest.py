import glob
import os

os.environ.setdefault("REVEAL_B200_TEST_HOOKS", "1")  # reveallib._load (library injection) only works with this set before the import
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


class _Impl(object):
    """One implementation of the drop-in surface: .index / .error (32-bit module) and .index64."""

    def __init__(self, name, mod32, mod64):
        self.name, self.index, self.error, self.index64, self.mod32, self.mod64 = name, mod32.index, mod32.error, mod64.index, mod32, mod64


@pytest.fixture
def emu_reveallib(emu_lib):
    """The drop-in `index` type (the compiled CPython extension reveal_b200.reveallib, csrc/ext/reveallib_module.cpp) with the
    emulated kernels injected in place of the CUDA library."""
    emu_path = os.path.join(ROOT, "tests", "emu", "_build", "libreveal_emu.so")
    from reveal_b200 import build
    build.build_extension()
    from reveal_b200 import reveallib, reveallib64
    reveallib._load(emu_path)
    reveallib64._load(emu_path)
    try:
        yield _Impl("ext", reveallib, reveallib64)
    finally:
        product = os.path.join(ROOT, "reveal_b200", "libreveal_b200.so")
        if os.path.exists(product):
            reveallib._load(product)
            reveallib64._load(product)


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library on a real GPU (gpu-marked tests only)."""
    from reveal_b200 import _native
    return _native.lib()
