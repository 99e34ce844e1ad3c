"""Batched mirror of simd/keccakf1600 (f1600x.go:30-131) and of the one-shot functions of internal/sha3
(hashes.go:21-60, shake.go:56-110) over the C ABI.

The reference permutes 4 interleaved states per call (StateX4); here a call permutes n independent states,
one per GPU thread.  numpy arrays (host: staged through the GPU inside the call) or torch CUDA tensors (in place /
asynchronous on the current torch stream).
"""
from __future__ import annotations

import numpy as np

from ._ffi import check, lib


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _sync_stream(x) -> None:
    if _is_torch(x):
        import torch
        check(lib().cb200_set_stream(torch.cuda.current_stream().cuda_stream))


def _ptr(x) -> int:
    if _is_torch(x):
        assert x.is_cuda and x.is_contiguous()
        return x.data_ptr()
    assert isinstance(x, np.ndarray) and x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def permute_(states, turbo: bool = False):
    """(*StateX4).Permute / KeccakF1600 on every state of the batch, in place: states (n, 25) uint64 (int64 for torch)."""
    n = (states.numel() if _is_torch(states) else states.size) // 25
    assert states.dtype.itemsize == 8 if not _is_torch(states) else states.element_size() == 8
    _sync_stream(states)
    check(lib().cb200_keccak_f1600(_ptr(states), n, 1 if turbo else 0))
    return states


def _sponge(bits: int, msgs, outlen: int):
    if _is_torch(msgs):
        import torch
        n, inlen = msgs.shape
        out = torch.empty((n, outlen), dtype=torch.uint8, device=msgs.device)
    else:
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        n, inlen = msgs.shape
        out = np.empty((n, outlen), dtype=np.uint8)
    _sync_stream(msgs)
    check(lib().cb200_sha3(bits, _ptr(msgs) if inlen else None, inlen, inlen, _ptr(out), outlen, n))
    return out


def shake128(msgs, outlen: int):
    """ShakeSum128 of n equal-length messages: (n, inlen) uint8 -> (n, outlen) uint8."""
    return _sponge(128, msgs, outlen)


def shake256(msgs, outlen: int):
    return _sponge(256, msgs, outlen)


def sha3_256(msgs):
    return _sponge(-256, msgs, 32)


def sha3_512(msgs):
    return _sponge(-512, msgs, 64)
