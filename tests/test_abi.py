"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/free_hunch_b200.h declares
(no compute calls here); the ctypes binding and the header agree; entry points refuse to work without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "free_hunch_b200.h")
LIB = os.path.join(ROOT, "free_hunch_b200", "libfreehunch_b200.so")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fh_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return ctypes.CDLL(LIB)


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_matches_header():
    from free_hunch_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_symbols()


def test_header_cites_the_reference():
    src = open(HEADER).read()
    for ref in ("online_update_bfgs.py", "cg.py", "conditioning_mechanisms.py", "measurements.py", "utils_sisr.py"):
        assert ref in src


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(lib):
    lib.fh_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    rc = lib.fh_ctx_create(ctypes.byref(h), 0, 1, 3, 256, 256, 64)
    assert rc == -2                                     # FH_ECUDA
    assert b"cuda" in lib.fh_last_error().lower()
    from free_hunch_b200 import engine, _lib
    with pytest.raises(_lib.FreeHunchLibraryError):
        engine.Context("cpu")
