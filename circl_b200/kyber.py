"""Batched mirror of the reference's Kyber ``Poly`` method surface
(pke/kyber/internal/common/generic.go:7-77, poly.go) over the C ABI.

Arguments are either numpy int16 arrays (host memory: copied to the GPU and
back inside the call) or torch CUDA int16 tensors (used in place, asynchronous
on the current torch stream).  Shapes are (..., 256).  Standard coefficient
order; bit-identical to the reference's *Generic functions.
"""
from __future__ import annotations

import numpy as np

from ._ffi import check, lib

N = 256
Q = 3329
OP_ADD, OP_SUB, OP_BARRETT, OP_NORMALIZE, OP_TOMONT = range(5)


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> int:
    if _is_torch(x):
        assert x.is_cuda and x.is_contiguous() and x.element_size() == 2
        return x.data_ptr()
    assert isinstance(x, np.ndarray) and x.dtype == np.int16 and x.flags["C_CONTIGUOUS"]
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
    """(*Poly).NTT on every polynomial of the batch, in place (generic.go:24)."""
    _sync_stream(p)
    check(lib().cb200_kyber_ntt(_ptr(p), _numel(p) // N, 0))
    return p


def inv_ntt_(p):
    """(*Poly).InvNTT, in place (generic.go:36)."""
    _sync_stream(p)
    check(lib().cb200_kyber_ntt(_ptr(p), _numel(p) // N, 1))
    return p


def mul_hat(a, b, out=None):
    """(*Poly).MulHat (generic.go:49)."""
    out = _like(a) if out is None else out
    _sync_stream(a)
    check(lib().cb200_kyber_mulhat(_ptr(out), _ptr(a), _ptr(b), _numel(a) // N))
    return out


def poly_dot_hat(a, b, k: int):
    """PolyDotHat over a batch: a, b (n, k, 256) -> (n, 256)  (vec.go:30-37)."""
    n = _numel(a) // (N * k)
    if _is_torch(a):
        import torch
        out = torch.empty((n, N), dtype=a.dtype, device=a.device)
    else:
        out = np.empty((n, N), dtype=np.int16)
    _sync_stream(a)
    check(lib().cb200_kyber_dot(_ptr(out), _ptr(a), _ptr(b), k, n))
    return out


def _unary(op, p, out=None):
    out = _like(p) if out is None else out
    _sync_stream(p)
    check(lib().cb200_kyber_poly_op(op, _ptr(out), _ptr(p), None, _numel(p) // N))
    return out


def _binary(op, a, b, out=None):
    out = _like(a) if out is None else out
    _sync_stream(a)
    check(lib().cb200_kyber_poly_op(op, _ptr(out), _ptr(a), _ptr(b), _numel(a) // N))
    return out


def add(a, b, out=None):
    return _binary(OP_ADD, a, b, out)


def sub(a, b, out=None):
    return _binary(OP_SUB, a, b, out)


def barrett_reduce(p, out=None):
    return _unary(OP_BARRETT, p, out)


def normalize(p, out=None):
    return _unary(OP_NORMALIZE, p, out)


def to_mont(p, out=None):
    return _unary(OP_TOMONT, p, out)


# ---- samplers and serialisation (pke/kyber/internal/common/sample.go, poly.go) ----
def _u8(x) -> int:
    if _is_torch(x):
        assert x.is_cuda and x.is_contiguous() and x.element_size() == 1
        return x.data_ptr()
    assert isinstance(x, np.ndarray) and x.dtype == np.uint8 and x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def _out(ref, shape, dtype):
    if _is_torch(ref):
        import torch
        return torch.empty(shape, dtype={np.int16: torch.int16, np.uint8: torch.uint8}[dtype], device=ref.device)
    return np.empty(shape, dtype=dtype)


def derive_uniform(seeds, xy):
    """(*Poly).DeriveUniform for n (seed, x, y): seeds (n, 32) or (32,) shared, xy (n, 2) uint8 -> (n, 256) int16."""
    n = _numel(xy) // 2
    shared = _numel(seeds) == 32
    out = _out(xy, (n, N), np.int16)
    _sync_stream(xy)
    check(lib().cb200_kyber_derive_uniform(_ptr(out), _u8(seeds), 0 if shared else 32, _u8(xy), n))
    return out


def derive_noise(seeds, nonces, eta: int):
    """(*Poly).DeriveNoise(eta): seeds (n, 32) or (32,) shared, nonces (n,) uint8 -> (n, 256) int16."""
    n = _numel(nonces)
    shared = _numel(seeds) == 32
    out = _out(nonces, (n, N), np.int16)
    _sync_stream(nonces)
    check(lib().cb200_kyber_derive_noise(_ptr(out), eta, _u8(seeds), 0 if shared else 32, _u8(nonces), n))
    return out


def pack(p):
    """(*Poly).Pack: (n, 256) normalised int16 -> (n, 384) uint8."""
    n = _numel(p) // N
    out = _out(p, (n, 384), np.uint8)
    _sync_stream(p)
    check(lib().cb200_kyber_pack(_u8(out), _ptr(p), n))
    return out


def unpack(buf):
    """(*Poly).Unpack: (n, 384) uint8 -> (n, 256) int16."""
    n = _numel(buf) // 384
    out = _out(buf, (n, N), np.int16)
    _sync_stream(buf)
    check(lib().cb200_kyber_unpack(_ptr(out), _u8(buf), n))
    return out


def compress(p, d: int):
    """(*Poly).CompressTo (d in 4, 5, 10, 11) / CompressMessageTo (d = 1): (n, 256) normalised -> (n, 32 d) uint8."""
    n = _numel(p) // N
    out = _out(p, (n, 32 * d), np.uint8)
    _sync_stream(p)
    check(lib().cb200_kyber_compress(_u8(out), _ptr(p), d, n))
    return out


def decompress(buf, d: int):
    """(*Poly).Decompress / DecompressMessage: (n, 32 d) uint8 -> (n, 256) int16."""
    n = _numel(buf) // (32 * d)
    out = _out(buf, (n, N), np.int16)
    _sync_stream(buf)
    check(lib().cb200_kyber_decompress(_ptr(out), _u8(buf), d, n))
    return out
