"""
pytest configuration.

Two tiers:
  * not gpu : oracle vs golden vectors, host logic, C-ABI symbol check, and the
              kernels' own sources run through the CPU execution model
              (tests/emu) against the oracle at small sizes;
  * gpu     : the parity tests proper -- libsetk_b200.so (sm_100a) through the
              C-ABI against the oracle, the golden fixtures, and
              size-independent properties at BASELINE.json's full sizes.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libsetk_b200_emu.so")
CSRC = os.path.join(ROOT, "setk_b200", "csrc")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def _emu_stale():
    if not os.path.exists(EMU_LIB):
        return True
    t = os.path.getmtime(EMU_LIB)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    srcs += [os.path.join(EMU_DIR, "cuda_emu.h"), os.path.join(EMU_DIR, "cuda_emu.cc"),
             os.path.join(ROOT, "include", "setk_b200.h")]
    return any(os.path.getmtime(s) > t for s in srcs)


@pytest.fixture(scope="session")
def emu_library_path():
    if _emu_stale():
        subprocess.run([os.path.join(EMU_DIR, "build_emu.sh")], check=True,
                       stdout=subprocess.DEVNULL)
    return EMU_LIB


@pytest.fixture()
def emu(emu_library_path):
    """Bind the CPU-emulated kernels; yields the torch device to use."""
    import torch
    from setk_b200 import _lib
    _lib.use_library(emu_library_path)
    yield torch.device("cpu")


@pytest.fixture()
def cuda():
    """Bind the real library; yields cuda:0.  Fails (not skips) without the .so."""
    import torch
    from setk_b200 import _lib
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    assert os.path.exists(_lib.DEFAULT_LIBRARY), \
        "libsetk_b200.so missing: run __graft_entry__.build() first"
    _lib.use_library(_lib.DEFAULT_LIBRARY)
    yield torch.device("cuda:0")
