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


def test_every_scheme_fails_loudly_without_a_device():
    """No CPU fallback anywhere: the scheme layers raise the library's NOT_INIT error instead of computing."""
    import numpy as np
    import pytest
    from circl_b200 import _ffi, hybrid, mldsa, mlkem
    if _ffi.lib().cb200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    seeds64 = np.zeros((2, 64), dtype=np.uint8)
    seeds32 = np.zeros((2, 32), dtype=np.uint8)
    for call in (lambda: mlkem.ByName("ML-KEM-768").DeriveKeyPairBatch(seeds64),
                 lambda: mlkem.ByName("Kyber768").DeriveKeyPairBatch(seeds64),
                 lambda: mldsa.ByName("ML-DSA-65").DeriveKeyBatch(seeds32),
                 lambda: mldsa.ByName("Dilithium3").DeriveKeyBatch(seeds32),
                 lambda: hybrid.ByName("X-Wing").DeriveKeyPairBatch(seeds32),
                 lambda: hybrid.ByName("X25519MLKEM768").DeriveKeyPairBatch(seeds64),
                 lambda: hybrid.x25519_keygen(seeds32)):
        with pytest.raises(_ffi.Cb200Error) as ei:
            call()
        assert "not initialised" in str(ei.value)


def test_scheme_registries_match_the_reference_names():
    # kem/schemes/schemes.go:30-55 and sign/schemes/schemes.go:33-50, restricted to the lattice families and their callers
    from circl_b200 import hybrid, mldsa, mlkem
    assert [s.Name() for s in mlkem.All()] == ["ML-KEM-512", "ML-KEM-768", "ML-KEM-1024", "Kyber512", "Kyber768", "Kyber1024"]
    assert [s.Name() for s in mldsa.All()] == ["ML-DSA-44", "ML-DSA-65", "ML-DSA-87", "Dilithium2", "Dilithium3", "Dilithium5"]
    assert sorted(s.Name() for s in hybrid.All()) == ["Kyber512-X25519", "Kyber768-X25519", "X-Wing", "X25519MLKEM768"]
    assert mlkem.ByName("kyber768") is mlkem.ByName("Kyber768") and hybrid.ByName("nope") is None
    x = hybrid.ByName("X25519MLKEM768")
    assert (x.SeedSize(), x.EncapsulationSeedSize(), x.SharedKeySize()) == (64, 32, 64)  # hybrid.go:91-117
    assert (x.PublicKeySize(), x.PrivateKeySize(), x.CiphertextSize()) == (1216, 2432, 1120)
