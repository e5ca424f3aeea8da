"""CPU: the C-ABI shared library loads and exports every symbol include/romhc.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "romhc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(romhc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from romhighcontrast_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/romhc.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_version_and_error_string_without_gpu():
    from romhighcontrast_b200 import _lib
    lib = _lib.load()
    assert lib.romhc_version() >= 100
    assert isinstance(lib.romhc_last_error(), bytes)
    assert lib.romhc_launch_count() >= 0


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from romhighcontrast_b200 import _lib
    from romhighcontrast_b200.engine import Engine
    with pytest.raises(_lib.RomhcError):
        Engine((2, 2), 4)
    # the C entry point itself also refuses (no CPU fallback anywhere)
    h = ctypes.c_void_p()
    rc = _lib.load().romhc_create(2, 2, 4, 0, ctypes.byref(h))
    assert rc == _lib.ERR_CUDA and b"no CPU fallback" in _lib.load().romhc_last_error()
