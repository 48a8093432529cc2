"""
CPU tier: the C-ABI shared library loads (no GPU needed to load it) and exports
every symbol include/setk_b200.h declares; no compute calls are made here.
"""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "setk_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(setk_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_boundary():
    names = declared_symbols()
    for must in ("setk_plan_create", "setk_stft", "setk_stft_cov", "setk_cov", "setk_weights",
                 "setk_apply", "setk_istft", "setk_apply_istft", "setk_version"):
        assert must in names


def test_ctypes_binding_covers_header():
    from setk_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()


@pytest.mark.parametrize("which", ["product", "emu"])
def test_library_exports_every_declared_symbol(which, emu_library_path):
    from setk_b200 import _lib
    if which == "product":
        path = _lib.DEFAULT_LIBRARY
        import __graft_entry__ as g
        if g._stale():
            pytest.skip("libsetk_b200.so is not built for the current sources "
                        "(run __graft_entry__.build())")
    else:
        path = emu_library_path
    lib = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    lib.setk_version.restype = ctypes.c_int
    assert lib.setk_version() == 100


def test_product_loader_fails_loudly_without_library(tmp_path, monkeypatch):
    from setk_b200 import _lib
    monkeypatch.setattr(_lib, "_cdll", None)
    monkeypatch.setattr(_lib, "DEFAULT_LIBRARY", str(tmp_path / "libsetk_b200.so"))
    with pytest.raises(ImportError):
        _lib.library()
