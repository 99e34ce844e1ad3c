"""Batched mirror of the reference's Dilithium ``Poly`` method surface
(sign/internal/dilithium/generic.go:11-89, poly.go) over the C ABI.

numpy uint32 arrays (host: staged through the GPU inside the call) or torch CUDA
int32/uint32 tensors (in place, asynchronous on the current torch stream);
shapes (..., 256).  Bit-identical to the reference's *Generic functions.
"""
from __future__ import annotations

import numpy as np

from ._ffi import check, lib

N = 256
Q = 8380417
OP_ADD, OP_SUB, OP_REDUCE_LE2Q, OP_NORMALIZE, OP_NORMALIZE_LE2Q, OP_MUL_2D = range(6)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> int:
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_cuda and x.is_contiguous() and x.element_size() == 4
        return x.data_ptr()
    assert isinstance(x, np.ndarray) and x.dtype == np.uint32 and x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def _numel(x) -> int:
    return x.numel() if _is_torch(x) else x.size


def _sync_stream(x) -> None:
    if _is_torch(x):
        import torch
        check(lib().cb200_set_stream(torch.cuda.current_stream().cuda_stream))


def _like(x):
    if _is_torch(x):
        import torch
        return torch.empty_like(x)
    return np.empty_like(x)


def ntt_(p):
    _sync_stream(p)
    check(lib().cb200_dil_ntt(_ptr(p), _numel(p) // N, 0))
    return p


def inv_ntt_(p):
    _sync_stream(p)
    check(lib().cb200_dil_ntt(_ptr(p), _numel(p) // N, 1))
    return p


def mul_hat(a, b, out=None):
    out = _like(a) if out is None else out
    _sync_stream(a)
    check(lib().cb200_dil_mulhat(_ptr(out), _ptr(a), _ptr(b), _numel(a) // N))
    return out


def poly_dot_hat(a, b, k: int):
    """PolyDotHat: a, b (n, k, 256) -> (n, 256)  (mat.go:52-59)."""
    n = _numel(a) // (N * k)
    if _is_torch(a):
        import torch
        out = torch.empty((n, N), dtype=a.dtype, device=a.device)
    else:
        out = np.empty((n, N), dtype=np.uint32)
    _sync_stream(a)
    check(lib().cb200_dil_dot(_ptr(out), _ptr(a), _ptr(b), k, n))
    return out


def _op(op, a, b=None, out=None):
    out = _like(a) if out is None else out
    _sync_stream(a)
    check(lib().cb200_dil_poly_op(op, _ptr(out), _ptr(a), _ptr(b), _numel(a) // N))
    return out


def add(a, b, out=None):
    return _op(OP_ADD, a, b, out)


def sub(a, b, out=None):
    return _op(OP_SUB, a, b, out)


def reduce_le2q(p, out=None):
    return _op(OP_REDUCE_LE2Q, p, None, out)


def normalize(p, out=None):
    return _op(OP_NORMALIZE, p, None, out)


def normalize_assuming_le2q(p, out=None):
    return _op(OP_NORMALIZE_LE2Q, p, None, out)


def mul_by_2_to_d(p, out=None):
    return _op(OP_MUL_2D, p, None, out)


def exceeds(p, bound: int):
    """(*Poly).Exceeds per polynomial -> uint8 flags of shape (n,)."""
    n = _numel(p) // N
    if _is_torch(p):
        import torch
        flags = torch.empty((n,), dtype=torch.uint8, device=p.device)
        _sync_stream(p)
        check(lib().cb200_dil_exceeds(_ptr(p), bound, flags.data_ptr(), n))
        return flags
    flags = np.empty((n,), dtype=np.uint8)
    check(lib().cb200_dil_exceeds(_ptr(p), bound, flags.ctypes.data, n))
    return flags


# ---- samplers and packers (sign/mldsa/mldsa{44,65,87}/internal/sample.go, sign/internal/dilithium/generic.go) ----
def _raw(x) -> int:
    if _is_torch(x):
        assert x.is_cuda and x.is_contiguous()
        return x.data_ptr()
    assert isinstance(x, np.ndarray) and x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def _new(ref, shape, dtype):
    if _is_torch(ref):
        import torch
        return torch.empty(shape, dtype={np.uint32: torch.int32, np.uint8: torch.uint8}[dtype], device=ref.device)
    return np.empty(shape, dtype=dtype)


def derive_uniform(seeds, nonces):
    """PolyDeriveUniform: seeds (n, 32) uint8 or (32,) shared, nonces (n,) uint16 -> (n, 256) uint32."""
    n = _numel(nonces)
    shared = _numel(seeds) == 32
    out = _new(nonces, (n, N), np.uint32)
    _sync_stream(nonces)
    check(lib().cb200_dil_derive_uniform(_raw(out), _raw(seeds), 0 if shared else 32, _raw(nonces), n))
    return out


def derive_leq_eta(mode: int, seeds, nonces):
    """PolyDeriveUniformLeqEta of ML-DSA-`mode`: seeds (n, 64) or (64,) shared, nonces (n,) uint16."""
    n = _numel(nonces)
    shared = _numel(seeds) == 64
    out = _new(nonces, (n, N), np.uint32)
    _sync_stream(nonces)
    check(lib().cb200_dil_derive_leq_eta(mode, _raw(out), _raw(seeds), 0 if shared else 64, _raw(nonces), n))
    return out


def derive_le_gamma1(mode: int, seeds, nonces):
    """PolyDeriveUniformLeGamma1 of ML-DSA-`mode`: seeds (n, 64) or (64,) shared, nonces (n,) uint16."""
    n = _numel(nonces)
    shared = _numel(seeds) == 64
    out = _new(nonces, (n, N), np.uint32)
    _sync_stream(nonces)
    check(lib().cb200_dil_derive_le_gamma1(mode, _raw(out), _raw(seeds), 0 if shared else 64, _raw(nonces), n))
    return out


def derive_ball(mode: int, seeds):
    """PolyDeriveUniformBall of ML-DSA-`mode`: seeds (n, 32 | 48 | 64) uint8 (c~) -> (n, 256) uint32."""
    ln = {44: 32, 65: 48, 87: 64}[mode]
    n = _numel(seeds) // ln
    out = _new(seeds, (n, N), np.uint32)
    _sync_stream(seeds)
    check(lib().cb200_dil_derive_ball(mode, _raw(out), _raw(seeds), ln, n))
    return out


def power2round(p):
    """(*Poly).Power2Round: normalised (n, 256) -> (a0 + Q, a1)."""
    n = _numel(p) // N
    a0, a1 = _like(p), _like(p)
    _sync_stream(p)
    check(lib().cb200_dil_power2round(_ptr(a0), _ptr(a1), _ptr(p), n))
    return a0, a1


def pack_le16(p):
    """(*Poly).PackLe16: (n, 256) with coefficients < 16 -> (n, 128) uint8."""
    n = _numel(p) // N
    out = _new(p, (n, 128), np.uint8)
    _sync_stream(p)
    check(lib().cb200_dil_pack_le16(_raw(out), _ptr(p), n))
    return out
