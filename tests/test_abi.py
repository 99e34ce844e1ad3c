"""CPU-only: the C-ABI library loads and exports every symbol include/circl_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "circl_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cb200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from circl_b200 import _ffi
    L = ctypes.CDLL(_ffi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # and the Python binding table covers the same set
    assert sorted(_ffi.SYMBOLS) == names


def test_compute_fails_loudly_without_init():
    from circl_b200 import _ffi
    L = _ffi.lib()
    import numpy as np
    p = np.zeros((1, 256), dtype=np.int16)
    rc = L.cb200_kyber_ntt(p.ctypes.data, 1, 0)
    assert rc != 0 and b"not initialised" in L.cb200_last_error()
