"""sign.Scheme mirror for ML-DSA-44 / ML-DSA-65 / ML-DSA-87 over the C ABI.

Mirrors sign/sign.go:48-119 (sign.Scheme, SignatureOpts, errors) and
sign/mldsa/mldsa65/dilithium.go:256-366 for the path this repository accelerates
(private-key expansion + Sign) and adds ``SignBatch``.  Wrong key types raise
like the reference panics; an over-long context raises ErrContextTooLong.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import Cb200Error, check, lib

SIGN_INTERNAL = 1


class SignError(Exception):
    pass


class ErrContextNotSupported(SignError):   # sign.ErrContextNotSupported (round-3 Dilithium, mode3/dilithium.go:227-229)
    pass


class ErrContextTooLong(SignError):   # sign.ErrContextTooLong
    pass


class ErrPrivKeySize(SignError):
    pass


class SignatureOpts:
    """sign.SignatureOpts (sign/sign.go:14-18)."""

    def __init__(self, Context: bytes = b""):
        self.Context = Context


class PublicKey:
    def __init__(self, scheme, packed: bytes):
        self._scheme, self._packed = scheme, bytes(packed)

    def Scheme(self):
        return self._scheme

    def MarshalBinary(self) -> bytes:
        return self._packed

    def Equal(self, other) -> bool:
        return isinstance(other, PublicKey) and other._packed == self._packed


class ErrPubKeySize(SignError):
    pass


class PrivateKey:
    def __init__(self, scheme, packed: bytes):
        self._scheme, self._packed = scheme, bytes(packed)

    def Scheme(self):
        return self._scheme

    def MarshalBinary(self) -> bytes:
        return self._packed

    def Equal(self, other) -> bool:
        return isinstance(other, PrivateKey) and other._packed == self._packed


class Scheme:
    """sign.Scheme for one parameter set (sign/dilithium/gen.go:80-162): mode 44, 65, 87 = ML-DSA; mode 2, 3, 5 = the
    round-3 Dilithium2/3/5 of sign/dilithium/mode{2,3,5} (no context string, no rnd, 32-byte tr and c~)."""

    def __init__(self, name: str = "ML-DSA-65", mode: int = 65):
        self._name, self._mode = name, mode

    def Name(self) -> str:
        return self._name

    def PublicKeySize(self) -> int:
        return {44: 1312, 65: 1952, 87: 2592, 2: 1312, 3: 1952, 5: 2592}[self._mode]

    def PrivateKeySize(self) -> int:
        return {44: 2560, 65: 4032, 87: 4896, 2: 2528, 3: 4000, 5: 4864}[self._mode]

    def SignatureSize(self) -> int:
        return {44: 2420, 65: 3309, 87: 4627, 2: 2420, 3: 3293, 5: 4595}[self._mode]

    def SeedSize(self) -> int:
        return 32

    def SupportsContext(self) -> bool:
        return self._mode > 10  # sign/dilithium/mode3/dilithium.go:208-210: round 3 has no context

    def UnmarshalBinaryPrivateKey(self, buf: bytes) -> PrivateKey:
        if len(buf) != self.PrivateKeySize():
            raise ErrPrivKeySize("sign: invalid private key size")  # dilithium.go:346-349
        return PrivateKey(self, buf)

    def Sign(self, sk: PrivateKey, message: bytes, opts: SignatureOpts | None = None) -> bytes:
        """sign.Scheme.Sign (dilithium.go:282-303): deterministic, batch of one."""
        if not isinstance(sk, PrivateKey):
            raise TypeError("sign: wrong private key type")  # the reference panics with sign.ErrTypeMismatch
        ctx = opts.Context if opts is not None else b""
        return self.SignBatch(sk, [message], ctx=ctx)[0].tobytes()

    def SignBatch(self, sks, messages, ctx: bytes = b"", rnd=None, internal: bool = False, return_attempts=False):
        """Batched Sign.  sks: one PrivateKey (shared, expanded once) or an (n, 4032) uint8 array
        (expanded on the device per op).  messages: list of bytes.  rnd: None (deterministic) or
        (n, 32) uint8.  internal=True selects ML-DSA.Sign_internal (no context framing; ACVP)."""
        if ctx and not self.SupportsContext():
            raise ErrContextNotSupported("sign: context not supported")
        if len(ctx) > 255:
            raise ErrContextTooLong("sign: context string too long")
        n = len(messages)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in messages], dtype=np.uint64)
        blob = np.frombuffer(b"".join(messages) + b"\0" * 8, dtype=np.uint8)
        if isinstance(sks, PrivateKey):
            sk = np.frombuffer(sks._packed, dtype=np.uint8)
            stride = 0
        else:
            sk = np.ascontiguousarray(sks, dtype=np.uint8)
            if sk.shape != (n, self.PrivateKeySize()):
                raise ErrPrivKeySize("sign: invalid private key size")
            stride = self.PrivateKeySize()
        sig = np.empty((n, self.SignatureSize()), dtype=np.uint8)
        status = np.zeros((n,), dtype=np.uint8)
        attempts = C.c_uint64(0)
        r = None if rnd is None else np.ascontiguousarray(rnd, dtype=np.uint8)
        cbuf = (C.c_uint8 * max(1, len(ctx))).from_buffer_copy(ctx or b"\0")
        check(lib().cb200_mldsa_sign(self._mode, sk.ctypes.data, stride, blob.ctypes.data, off.ctypes.data,
                                       C.cast(cbuf, C.c_void_p), len(ctx), None if r is None else r.ctypes.data,
                                       sig.ctypes.data, status.ctypes.data, n, SIGN_INTERNAL if internal else 0,
                                       C.cast(C.pointer(attempts), C.c_void_p)))
        if return_attempts:
            return sig, int(attempts.value)
        return sig

    def DeriveKey(self, seed: bytes):
        """sign.Scheme.DeriveKey (dilithium.go:266-276): key pair from a 32-byte seed (batch of one)."""
        if len(seed) != self.SeedSize():
            raise ValueError("sign: invalid seed size")  # the reference panics with sign.ErrSeedSize
        pk, sk = self.DeriveKeyBatch(np.frombuffer(seed, dtype=np.uint8).reshape(1, 32))
        return PublicKey(self, pk[0].tobytes()), PrivateKey(self, sk[0].tobytes())

    def GenerateKey(self):
        import os
        return self.DeriveKey(os.urandom(32))

    def DeriveKeyBatch(self, seeds):
        """seeds: (n, 32) uint8 -> (pk (n, 1952), sk (n, 4032)) packed keys."""
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8)
        if seeds.ndim != 2 or seeds.shape[1] != 32:
            raise ValueError("sign: invalid seed size")
        n = seeds.shape[0]
        pk = np.empty((n, self.PublicKeySize()), dtype=np.uint8)
        sk = np.empty((n, self.PrivateKeySize()), dtype=np.uint8)
        check(lib().cb200_mldsa_keygen(self._mode, seeds.ctypes.data, pk.ctypes.data, sk.ctypes.data, n))
        return pk, sk

    def UnmarshalBinaryPublicKey(self, buf: bytes) -> PublicKey:
        if len(buf) != self.PublicKeySize():
            raise ErrPubKeySize("sign: invalid public key size")  # dilithium.go:337-340
        return PublicKey(self, buf)

    def Verify(self, pk: PublicKey, message: bytes, signature: bytes, opts: SignatureOpts | None = None) -> bool:
        """sign.Scheme.Verify (dilithium.go:305-330), batch of one."""
        if not isinstance(pk, PublicKey):
            raise TypeError("sign: wrong public key type")
        ctx = opts.Context if opts is not None else b""
        if ctx and not self.SupportsContext():
            raise ErrContextNotSupported("sign: context not supported")
        if len(ctx) > 255 or len(signature) != self.SignatureSize():
            return False  # dilithium.go:116-118, internal/dilithium.go:82-84
        return bool(self.VerifyBatch(pk, [message], np.frombuffer(signature, dtype=np.uint8).reshape(1, -1), ctx=ctx)[0])

    def VerifyBatch(self, pks, messages, sigs, ctx: bytes = b"", internal: bool = False):
        """pks: one PublicKey (shared) or (n, 1952) uint8; sigs: (n, 3309) uint8 -> (n,) bool array."""
        n = len(messages)
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(m) for m in messages], dtype=np.uint64)
        blob = np.frombuffer(b"".join(messages) + b"\0" * 8, dtype=np.uint8)
        if isinstance(pks, PublicKey):
            pk = np.frombuffer(pks._packed, dtype=np.uint8)
            stride = 0
        else:
            pk = np.ascontiguousarray(pks, dtype=np.uint8)
            if pk.shape != (n, self.PublicKeySize()):
                raise ErrPubKeySize("sign: invalid public key size")
            stride = self.PublicKeySize()
        sigs = np.ascontiguousarray(sigs, dtype=np.uint8)
        assert sigs.shape == (n, self.SignatureSize())
        ok = np.zeros((n,), dtype=np.uint8)
        cbuf = (C.c_uint8 * max(1, len(ctx))).from_buffer_copy(ctx or b"\0")
        check(lib().cb200_mldsa_verify(self._mode, pk.ctypes.data, stride, blob.ctypes.data, off.ctypes.data,
                                         C.cast(cbuf, C.c_void_p), len(ctx), sigs.ctypes.data, ok.ctypes.data, n,
                                         SIGN_INTERNAL if internal else 0))
        return ok.astype(bool)


_SCHEMES = {"ml-dsa-44": Scheme("ML-DSA-44", 44), "ml-dsa-65": Scheme("ML-DSA-65", 65), "ml-dsa-87": Scheme("ML-DSA-87", 87),
            "dilithium2": Scheme("Dilithium2", 2), "dilithium3": Scheme("Dilithium3", 3), "dilithium5": Scheme("Dilithium5", 5)}


def ByName(name: str):
    """sign/schemes/schemes.go:69 -- case-insensitive lookup; None if unknown."""
    return _SCHEMES.get(name.lower())


def All():
    return list(_SCHEMES.values())
