"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes view of ``oracle/liboracle.so``, the plain-C CPU restatement of the
reference's generic module-lattice path (see ``oracle/oracle.h``).  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package; nothing under
``circl_b200/`` does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _declare(_lib)
    return _lib


_u8p = C.POINTER(C.c_uint8)
_i16p = C.POINTER(C.c_int16)
_u32p = C.POINTER(C.c_uint32)


def _declare(L: C.CDLL) -> None:
    L.orc_kyber_mont_reduce.restype = C.c_int16
    L.orc_kyber_mont_reduce.argtypes = [C.c_int32]
    L.orc_kyber_barrett_reduce.restype = C.c_int16
    L.orc_kyber_barrett_reduce.argtypes = [C.c_int16]
    L.orc_kyber_csubq.restype = C.c_int16
    L.orc_kyber_csubq.argtypes = [C.c_int16]
    L.orc_kyber_zetas.restype = _i16p
    for name in ("orc_mlkem_ek_size", "orc_mlkem_dk_size", "orc_mlkem_ct_size"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [C.c_int]
    L.orc_mlkem_encaps_batch.restype = C.c_int
    L.orc_mlkem_encaps_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                         C.c_void_p, C.c_size_t, C.c_int]
    L.orc_kyber_ntt_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    L.orc_kyber_mulhat_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.orc_kyber_dot_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_size_t]
    if hasattr(L, "orc_dil_ntt_batch"):
        L.orc_dil_ntt_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.orc_dil_mulhat_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_mldsa65_sign_batch.restype = C.c_int
        L.orc_mldsa65_sign_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_size_t, C.c_int]


def _buf(b) -> C.Array:
    return (C.c_uint8 * len(b)).from_buffer_copy(bytes(b))


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ Keccak
def keccak_f1600(state25) -> list[int]:
    a = (C.c_uint64 * 25)(*state25)
    lib().orc_keccak_f1600(a)
    return list(a)


def keccak_f1600_turbo(state25) -> list[int]:
    a = (C.c_uint64 * 25)(*state25)
    lib().orc_keccak_f1600_turbo(a)
    return list(a)


def _hash(fn, outlen: int, data: bytes) -> bytes:
    out = (C.c_uint8 * outlen)()
    if fn in ("orc_sha3_256", "orc_sha3_512"):
        getattr(lib(), fn)(out, _buf(data) if data else None, C.c_size_t(len(data)))
    else:
        getattr(lib(), fn)(out, C.c_size_t(outlen), _buf(data) if data else None, C.c_size_t(len(data)))
    return bytes(out)


def sha3_256(data: bytes) -> bytes:
    return _hash("orc_sha3_256", 32, data)


def sha3_512(data: bytes) -> bytes:
    return _hash("orc_sha3_512", 64, data)


def shake128(data: bytes, outlen: int) -> bytes:
    return _hash("orc_shake128", outlen, data)


def shake256(data: bytes, outlen: int) -> bytes:
    return _hash("orc_shake256", outlen, data)


# ------------------------------------------------------------------ Kyber polys
def kyber_zetas() -> np.ndarray:
    return np.ctypeslib.as_array(lib().orc_kyber_zetas(), shape=(128,)).copy()


def _poly_unary(fn: str, p: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(p, dtype=np.int16).copy()
    flat = q.reshape(-1, 256)
    f = getattr(lib(), fn)
    for row in flat:
        f(_ptr(row))
    return q


def kyber_ntt(p):
    q = np.ascontiguousarray(p, dtype=np.int16).copy()
    lib().orc_kyber_ntt_batch(_ptr(q), q.size // 256, 0)
    return q


def kyber_ntt_inplace_mt(p: np.ndarray, inverse: bool, nthreads: int) -> None:
    """In-place multi-threaded batch NTT (bench.py cpu baseline)."""
    assert p.dtype == np.int16 and p.flags["C_CONTIGUOUS"]
    lib().orc_kyber_ntt_batch_mt(_ptr(p), C.c_size_t(p.size // 256), int(inverse), int(nthreads))


def kyber_invntt(p):
    q = np.ascontiguousarray(p, dtype=np.int16).copy()
    lib().orc_kyber_ntt_batch(_ptr(q), q.size // 256, 1)
    return q


def kyber_mulhat(a, b):
    a = np.ascontiguousarray(a, dtype=np.int16)
    b = np.ascontiguousarray(b, dtype=np.int16)
    out = np.empty_like(a)
    lib().orc_kyber_mulhat_batch(_ptr(out), _ptr(a), _ptr(b), a.size // 256)
    return out


def kyber_dot(a, b, k: int):
    """a, b: (n, k, 256) -> (n, 256) PolyDotHat."""
    a = np.ascontiguousarray(a, dtype=np.int16)
    b = np.ascontiguousarray(b, dtype=np.int16)
    n = a.size // (256 * k)
    out = np.empty((n, 256), dtype=np.int16)
    lib().orc_kyber_dot_batch(_ptr(out), _ptr(a), _ptr(b), k, n)
    return out


def kyber_barrett(p):
    return _poly_unary("orc_kyber_barrett", p)


def kyber_normalize(p):
    return _poly_unary("orc_kyber_normalize", p)


def kyber_tomont(p):
    return _poly_unary("orc_kyber_tomont", p)


def kyber_pack(p) -> bytes:
    out = (C.c_uint8 * 384)()
    lib().orc_kyber_pack(out, _ptr(np.ascontiguousarray(p, dtype=np.int16)))
    return bytes(out)


def kyber_unpack(buf: bytes) -> np.ndarray:
    p = np.empty(256, dtype=np.int16)
    lib().orc_kyber_unpack(_ptr(p), _buf(buf))
    return p


def kyber_compress(p, d: int) -> bytes:
    out = (C.c_uint8 * (32 * d))()
    lib().orc_kyber_compress(out, _ptr(np.ascontiguousarray(p, dtype=np.int16)), d)
    return bytes(out)


def kyber_decompress(buf: bytes, d: int) -> np.ndarray:
    p = np.empty(256, dtype=np.int16)
    lib().orc_kyber_decompress(_ptr(p), _buf(buf), d)
    return p


def kyber_derive_noise(seed: bytes, nonce: int, eta: int) -> np.ndarray:
    p = np.empty(256, dtype=np.int16)
    lib().orc_kyber_derive_noise(_ptr(p), _buf(seed), C.c_size_t(len(seed)), C.c_uint8(nonce), eta)
    return p


def kyber_derive_uniform(seed: bytes, x: int, y: int) -> np.ndarray:
    p = np.empty(256, dtype=np.int16)
    lib().orc_kyber_derive_uniform(_ptr(p), _buf(seed), C.c_uint8(x), C.c_uint8(y))
    return p


# ------------------------------------------------------------------ ML-KEM
def mlkem_sizes(k: int):
    L = lib()
    return L.orc_mlkem_ek_size(k), L.orc_mlkem_dk_size(k), L.orc_mlkem_ct_size(k)


def mlkem_keygen(k: int, seed64: bytes):
    eksz, dksz, _ = mlkem_sizes(k)
    ek = (C.c_uint8 * eksz)()
    dk = (C.c_uint8 * dksz)()
    lib().orc_mlkem_keygen(k, ek, dk, _buf(seed64))
    return bytes(ek), bytes(dk)


def mlkem_encaps(k: int, ek: bytes, m: bytes):
    _, _, ctsz = mlkem_sizes(k)
    ct = (C.c_uint8 * ctsz)()
    ss = (C.c_uint8 * 32)()
    rc = lib().orc_mlkem_encaps(k, ct, ss, _buf(ek), _buf(m))
    if rc:
        raise ValueError("kem.ErrPubKey")
    return bytes(ct), bytes(ss)


def mlkem_decaps(k: int, dk: bytes, ct: bytes) -> bytes:
    ss = (C.c_uint8 * 32)()
    rc = lib().orc_mlkem_decaps(k, ss, _buf(dk), _buf(ct))
    if rc:
        raise ValueError("kem.ErrPrivKey")
    return bytes(ss)


def mlkem_encaps_batch(k: int, ek: np.ndarray, m: np.ndarray, nthreads: int = 1):
    """ek: (n, eksz) or (eksz,) uint8 (shared); m: (n, 32) uint8."""
    _, _, ctsz = mlkem_sizes(k)
    m = np.ascontiguousarray(m, dtype=np.uint8)
    ek = np.ascontiguousarray(ek, dtype=np.uint8)
    n = m.shape[0]
    stride = 0 if ek.ndim == 1 else ek.shape[1]
    ct = np.empty((n, ctsz), dtype=np.uint8)
    ss = np.empty((n, 32), dtype=np.uint8)
    fails = lib().orc_mlkem_encaps_batch(k, _ptr(ct), _ptr(ss), _ptr(ek), stride, _ptr(m), n, nthreads)
    return ct, ss, fails


def mlkem_encaps_batch_avx2(k: int, ek: np.ndarray, m: np.ndarray, nthreads: int = 1):
    """The AVX2 arm (kyber_avx2.c): same contract and same bytes as mlkem_encaps_batch."""
    _, _, ctsz = mlkem_sizes(k)
    n = m.shape[0]
    ek = np.ascontiguousarray(ek, dtype=np.uint8)
    ct = np.empty((n, ctsz), dtype=np.uint8)
    ss = np.empty((n, 32), dtype=np.uint8)
    L = lib()
    L.orc_mlkem_encaps_batch_avx2.restype = C.c_int
    fails = L.orc_mlkem_encaps_batch_avx2(C.c_int(k), _ptr(ct), _ptr(ss), _ptr(ek), C.c_size_t(ek.shape[1] if ek.ndim == 2 else 0),
                                          _ptr(np.ascontiguousarray(m)), C.c_size_t(n), C.c_int(nthreads))
    return ct, ss, int(fails)


def keccak_f1600_x4(states4x25) -> np.ndarray:
    """Four states (4, 25) uint64 -> permuted, through the interleaved StateX4 layout."""
    a = np.ascontiguousarray(np.asarray(states4x25, dtype=np.uint64).T.reshape(100))
    lib().orc_keccak_f1600_x4(_ptr(a))
    return a.reshape(25, 4).T.copy()


def mlkem_parse_keys(k: int, ek: np.ndarray) -> np.ndarray:
    """UnmarshalBinaryPublicKey for every row of ek: the cached fields (th, aT, hpk) of each key, (n, parsed_size) uint8."""
    L = lib()
    L.orc_mlkem_parsed_size.restype = C.c_size_t
    psz = int(L.orc_mlkem_parsed_size())
    ek = np.ascontiguousarray(ek, dtype=np.uint8)
    out = np.empty((ek.shape[0], psz), dtype=np.uint8)
    for i in range(ek.shape[0]):
        rc = L.orc_mlkem_pk_parse(C.c_int(k), C.c_void_p(out[i].ctypes.data), C.c_void_p(ek[i].ctypes.data))
        assert rc == 0
    return out


def mlkem_encaps_parsed_batch(k: int, parsed: np.ndarray, m: np.ndarray, nthreads: int = 1):
    """EncapsulateTo on pre-parsed keys (one per op)."""
    _, _, ctsz = mlkem_sizes(k)
    n = m.shape[0]
    ct = np.empty((n, ctsz), dtype=np.uint8)
    ss = np.empty((n, 32), dtype=np.uint8)
    lib().orc_mlkem_encaps_parsed_batch(C.c_int(k), _ptr(ct), _ptr(ss), _ptr(parsed), C.c_size_t(parsed.shape[1]),
                                        _ptr(np.ascontiguousarray(m)), C.c_size_t(n), C.c_int(nthreads))
    return ct, ss


# ------------------------------------------------------------------ Dilithium / ML-DSA-65
def dil_zetas():
    L = lib()
    L.orc_dil_zetas.restype = _u32p
    L.orc_dil_inv_zetas.restype = _u32p
    return (np.ctypeslib.as_array(L.orc_dil_zetas(), shape=(256,)).copy(),
            np.ctypeslib.as_array(L.orc_dil_inv_zetas(), shape=(256,)).copy())


def dil_ntt(p):
    q = np.ascontiguousarray(p, dtype=np.uint32).copy()
    lib().orc_dil_ntt_batch(_ptr(q), q.size // 256, 0)
    return q


def dil_invntt(p):
    q = np.ascontiguousarray(p, dtype=np.uint32).copy()
    lib().orc_dil_ntt_batch(_ptr(q), q.size // 256, 1)
    return q


def dil_mulhat(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    b = np.ascontiguousarray(b, dtype=np.uint32)
    out = np.empty_like(a)
    lib().orc_dil_mulhat_batch(_ptr(out), _ptr(a), _ptr(b), a.size // 256)
    return out


def dil_poly_op(op: int, a, b=None):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    out = np.empty_like(a)
    fa, fo = a.reshape(-1, 256), out.reshape(-1, 256)
    fb = None if b is None else np.ascontiguousarray(b, dtype=np.uint32).reshape(-1, 256)
    for i in range(fa.shape[0]):
        lib().orc_dil_poly_op(op, _ptr(fo[i]), _ptr(fa[i]), None if fb is None else _ptr(fb[i]))
    return out


def dil_exceeds(p, bound: int) -> bool:
    return bool(lib().orc_dil_exceeds(_ptr(np.ascontiguousarray(p, dtype=np.uint32)), C.c_uint32(bound)))


def _dil_sample(fn, seed: bytes, nonce=None):
    p = np.empty(256, dtype=np.uint32)
    if nonce is None:
        getattr(lib(), fn)(_ptr(p), _buf(seed))
    else:
        getattr(lib(), fn)(_ptr(p), _buf(seed), C.c_uint16(nonce))
    return p


def dil_derive_uniform(seed32, nonce):
    return _dil_sample("orc_dil_derive_uniform", seed32, nonce)


def dil_derive_leqeta(seed64, nonce):
    return _dil_sample("orc_dil_derive_leqeta", seed64, nonce)


def dil_derive_legamma1(seed64, nonce):
    return _dil_sample("orc_dil_derive_legamma1", seed64, nonce)


def dil_derive_ball(seed48):
    return _dil_sample("orc_dil_derive_ball", seed48)


def mldsa_derive_leqeta(mode: int, seed64, nonce):
    p = np.empty(256, dtype=np.uint32)
    lib().orc_mldsa_derive_leqeta(C.c_int(mode), _ptr(p), _buf(seed64), C.c_uint16(nonce))
    return p


def mldsa_derive_legamma1(mode: int, seed64, nonce):
    p = np.empty(256, dtype=np.uint32)
    lib().orc_mldsa_derive_legamma1(C.c_int(mode), _ptr(p), _buf(seed64), C.c_uint16(nonce))
    return p


def mldsa_derive_ball(mode: int, ctilde):
    p = np.empty(256, dtype=np.uint32)
    lib().orc_mldsa_derive_ball(C.c_int(mode), _ptr(p), _buf(ctilde))
    return p


def dil_power2round(p):
    p = np.ascontiguousarray(p, dtype=np.uint32)
    a0, a1 = np.empty(256, dtype=np.uint32), np.empty(256, dtype=np.uint32)
    lib().orc_dil_power2round(_ptr(p), _ptr(a0), _ptr(a1))
    return a0, a1


def dil_pack_le16(p) -> bytes:
    out = (C.c_uint8 * 128)()
    lib().orc_dil_pack_le16(out, _ptr(np.ascontiguousarray(p, dtype=np.uint32)))
    return bytes(out)


def mldsa65_keygen(seed32: bytes):
    pk = (C.c_uint8 * 1952)()
    sk = (C.c_uint8 * 4032)()
    lib().orc_mldsa65_keygen(pk, sk, _buf(seed32))
    return bytes(pk), bytes(sk)


def mldsa65_sign(sk: bytes, msg: bytes, ctx: bytes = b"", rnd: bytes = bytes(32), internal: bool = False):
    """Returns (signature, attempts)."""
    sig = (C.c_uint8 * 3309)()
    L = lib()
    L.orc_mldsa65_sign.restype = C.c_int
    n = L.orc_mldsa65_sign(sig, _buf(sk), _buf(msg) if msg else None, C.c_size_t(len(msg)),
                           _buf(ctx) if ctx else None, C.c_size_t(len(ctx)), _buf(rnd), int(internal))
    if n < 0:
        raise RuntimeError("sign: 576 attempts exhausted")
    return bytes(sig), n


def mldsa65_verify(pk: bytes, msg: bytes, sig: bytes, ctx: bytes = b"", internal: bool = False) -> bool:
    L = lib()
    L.orc_mldsa65_verify.restype = C.c_int
    return bool(L.orc_mldsa65_verify(_buf(pk), _buf(msg) if msg else None, C.c_size_t(len(msg)),
                                     _buf(ctx) if ctx else None, C.c_size_t(len(ctx)), _buf(sig),
                                     C.c_size_t(len(sig)), int(internal)))


def mldsa65_sign_batch(sk: np.ndarray, msgs: list[bytes], rnd=None, nthreads: int = 1):
    """sk: (4032,) shared or (n, 4032); external interface, empty ctx. Returns (sigs (n,3309), total attempts)."""
    n = len(msgs)
    sk = np.ascontiguousarray(sk, dtype=np.uint8)
    stride = 0 if sk.ndim == 1 else sk.shape[1]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    blob = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8)
    sig = np.empty((n, 3309), dtype=np.uint8)
    r = None if rnd is None else np.ascontiguousarray(rnd, dtype=np.uint8)
    att = lib().orc_mldsa65_sign_batch(_ptr(sig), _ptr(sk), stride, _ptr(blob), _ptr(off),
                                       None if r is None else _ptr(r), n, nthreads)
    if att < 0:
        raise RuntimeError("sign_batch failed")
    return sig, att


# ------------------------------------------------------------------ ML-DSA, run-time parameter set (44 / 65 / 87)
def mldsa_sizes(mode: int):
    L = lib()
    for f in ("orc_mldsa_pk_size", "orc_mldsa_sk_size", "orc_mldsa_sig_size"):
        getattr(L, f).restype = C.c_size_t
    return L.orc_mldsa_pk_size(mode), L.orc_mldsa_sk_size(mode), L.orc_mldsa_sig_size(mode)


def mldsa_keygen(mode: int, seed32: bytes):
    pksz, sksz, _ = mldsa_sizes(mode)
    pk, sk = (C.c_uint8 * pksz)(), (C.c_uint8 * sksz)()
    lib().orc_mldsa_keygen(mode, pk, sk, _buf(seed32))
    return bytes(pk), bytes(sk)


def mldsa_sign(mode: int, sk: bytes, msg: bytes, ctx: bytes = b"", rnd: bytes = bytes(32), internal: bool = False):
    _, _, sigsz = mldsa_sizes(mode)
    sig = (C.c_uint8 * sigsz)()
    L = lib()
    L.orc_mldsa_sign.restype = C.c_int
    n = L.orc_mldsa_sign(mode, sig, _buf(sk), _buf(msg) if msg else None, C.c_size_t(len(msg)),
                         _buf(ctx) if ctx else None, C.c_size_t(len(ctx)), _buf(rnd), int(internal))
    if n < 0:
        raise RuntimeError("sign: 576 attempts exhausted")
    return bytes(sig), n


def mldsa_verify(mode: int, pk: bytes, msg: bytes, sig: bytes, ctx: bytes = b"", internal: bool = False) -> bool:
    L = lib()
    L.orc_mldsa_verify.restype = C.c_int
    return bool(L.orc_mldsa_verify(mode, _buf(pk), _buf(msg) if msg else None, C.c_size_t(len(msg)),
                                   _buf(ctx) if ctx else None, C.c_size_t(len(ctx)), _buf(sig),
                                   C.c_size_t(len(sig)), int(internal)))


def mldsa_sign_batch(mode: int, sk: np.ndarray, msgs: list[bytes], rnd=None, nthreads: int = 1):
    n = len(msgs)
    _, sksz, sigsz = mldsa_sizes(mode)
    sk = np.ascontiguousarray(sk, dtype=np.uint8)
    stride = 0 if sk.ndim == 1 else sk.shape[1]
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(m) for m in msgs])
    blob = np.frombuffer(b"".join(msgs) or b"\0", dtype=np.uint8)
    sig = np.empty((n, sigsz), dtype=np.uint8)
    r = None if rnd is None else np.ascontiguousarray(rnd, dtype=np.uint8)
    L = lib()
    L.orc_mldsa_sign_batch.restype = C.c_int
    L.orc_mldsa_sign_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_size_t, C.c_int]
    att = L.orc_mldsa_sign_batch(mode, _ptr(sig), _ptr(sk), stride, _ptr(blob), _ptr(off),
                                 None if r is None else _ptr(r), n, nthreads)
    if att < 0:
        raise RuntimeError("sign_batch failed")
    return sig, att


# ------------------------------------------------------------------ round-3 Kyber KEM (kem/kyber)
def kyber_kem_keygen(k: int, seed64: bytes):
    eksz, dksz, _ = mlkem_sizes(k)
    ek, dk = (C.c_uint8 * eksz)(), (C.c_uint8 * dksz)()
    lib().orc_kyber_kem_keygen(k, ek, dk, _buf(seed64))
    return bytes(ek), bytes(dk)


def kyber_kem_encaps(k: int, ek: bytes, seed32: bytes):
    _, _, ctsz = mlkem_sizes(k)
    ct, ss = (C.c_uint8 * ctsz)(), (C.c_uint8 * 32)()
    lib().orc_kyber_kem_encaps(k, ct, ss, _buf(ek), _buf(seed32))
    return bytes(ct), bytes(ss)


def kyber_kem_decaps(k: int, dk: bytes, ct: bytes) -> bytes:
    ss = (C.c_uint8 * 32)()
    lib().orc_kyber_kem_decaps(k, ss, _buf(dk), _buf(ct))
    return bytes(ss)


# ------------------------------------------------------------------ X25519, X-Wing, kem/hybrid (hybrid.c)
HYBRID_IDS = {"X25519MLKEM768": 0, "Kyber768-X25519": 1, "Kyber512-X25519": 2}


def x25519(scalar: bytes, point: bytes | None = None):
    """(out, ok): dh/x25519 KeyGen (point None) or Shared; ok False for the low-order points."""
    out = (C.c_uint8 * 32)()
    ok = lib().orc_x25519(out, _buf(scalar), _buf(point) if point is not None else None)
    return bytes(out), bool(ok)


def xwing_keygen(seed32: bytes) -> bytes:
    pk = (C.c_uint8 * 1216)()
    lib().orc_xwing_keygen(pk, _buf(seed32))
    return bytes(pk)


def xwing_encaps(pk: bytes, eseed64: bytes):
    ct, ss = (C.c_uint8 * 1120)(), (C.c_uint8 * 32)()
    if lib().orc_xwing_encaps(ct, ss, _buf(pk), _buf(eseed64)):
        raise ValueError("kem.ErrPubKey")
    return bytes(ct), bytes(ss)


def xwing_decaps(sk32: bytes, ct: bytes) -> bytes:
    ss = (C.c_uint8 * 32)()
    lib().orc_xwing_decaps(ss, _buf(sk32), _buf(ct))
    return bytes(ss)


def hybrid_sizes(name: str):
    L, i = lib(), HYBRID_IDS[name]
    for f in (L.orc_hybrid_pk_size, L.orc_hybrid_sk_size, L.orc_hybrid_ct_size):
        f.restype = C.c_size_t
    return L.orc_hybrid_pk_size(i), L.orc_hybrid_sk_size(i), L.orc_hybrid_ct_size(i)


def hybrid_keygen(name: str, seed64: bytes):
    pksz, sksz, _ = hybrid_sizes(name)
    pk, sk = (C.c_uint8 * pksz)(), (C.c_uint8 * sksz)()
    lib().orc_hybrid_keygen(HYBRID_IDS[name], pk, sk, _buf(seed64))
    return bytes(pk), bytes(sk)


def hybrid_encaps(name: str, pk: bytes, seed32: bytes):
    """(ct, ss, rc): rc 1 = kem.ErrPubKey."""
    _, _, ctsz = hybrid_sizes(name)
    ct, ss = (C.c_uint8 * ctsz)(), (C.c_uint8 * 64)()
    rc = lib().orc_hybrid_encaps(HYBRID_IDS[name], ct, ss, _buf(pk), _buf(seed32))
    return bytes(ct), bytes(ss), rc


def hybrid_decaps(name: str, sk: bytes, ct: bytes):
    """(ss, rc): rc 1 = kem.ErrPubKey (low-order X25519 share), 2 = kem.ErrPrivKey."""
    ss = (C.c_uint8 * 64)()
    rc = lib().orc_hybrid_decaps(HYBRID_IDS[name], ss, _buf(sk), _buf(ct))
    return bytes(ss), rc
