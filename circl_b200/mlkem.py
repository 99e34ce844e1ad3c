"""kem.Scheme mirror for ML-KEM-512/768/1024 and round-3 Kyber512/768/1024 over the C ABI.

Mirrors the method names, argument meaning and error behaviour of
  kem/kem.go:33-121                      (kem.Scheme, kem.Err*)
  kem/mlkem/mlkem768/kyber.go:267-407    (scheme boilerplate)
for the path this repository accelerates (public-key parsing + Encapsulate),
and adds the batch entry points a per-op GPU call cannot do without
(SURVEY.md 8(b)).  Programmer errors (wrong lengths) raise like the
reference panics; data errors raise the kem.Err* mirror classes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._ffi import Cb200Error, check, lib


class KemError(Exception):
    pass


class ErrPubKeySize(KemError):       # kem.ErrPubKeySize, kem/kem.go:112
    pass


class ErrPubKey(KemError):           # kem.ErrPubKey, kem/kem.go:118 (ek not reduced mod q)
    pass


class ErrSeedSize(KemError):         # kem.ErrSeedSize
    pass


class ErrPrivKeySize(KemError):      # kem.ErrPrivKeySize
    pass


class ErrPrivKey(KemError):          # kem.ErrPrivKey (H(ek) in dk mismatches, kyber.go:226-228)
    pass


class ErrCiphertextSize(KemError):   # kem.ErrCiphertextSize
    pass


class ErrTypeMismatch(KemError):     # kem.ErrTypeMismatch
    pass


class PublicKey:
    """kem.PublicKey: holds the packed encapsulation key (immutable)."""

    def __init__(self, scheme: "Scheme", packed: bytes):
        self._scheme = scheme
        self._packed = bytes(packed)

    def Scheme(self):
        return self._scheme

    def MarshalBinary(self) -> bytes:
        return self._packed

    def Equal(self, other) -> bool:
        return isinstance(other, PublicKey) and other._scheme is self._scheme and other._packed == self._packed


class PrivateKey:
    """kem.PrivateKey: holds the packed decapsulation key (immutable)."""

    def __init__(self, scheme: "Scheme", packed: bytes):
        self._scheme = scheme
        self._packed = bytes(packed)

    def Scheme(self):
        return self._scheme

    def MarshalBinary(self) -> bytes:
        return self._packed

    def Equal(self, other) -> bool:
        return isinstance(other, PrivateKey) and other._scheme is self._scheme and other._packed == self._packed

    def Public(self) -> PublicKey:
        k = self._scheme._k
        return PublicKey(self._scheme, self._packed[384 * k:384 * k + 384 * k + 32])


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class Scheme:
    def __init__(self, name: str, k: int, round3: bool = False):
        self._name, self._k, self._round3 = name, k, round3

    # One place decides which C entry points serve this scheme: FIPS 203 (kem/mlkem) or the round-3
    # submission (kem/kyber/kyber768/kyber.go), whose calls carry no status array because round-3
    # key parsing cannot fail (no modulus check, no H(ek) check).
    def _c_encaps(self, ek, stride, seeds, ct, ss, status, n):
        if self._round3:
            return lib().cb200_kyber_kem_encaps(self._k, ek, stride, seeds, ct, ss, n)
        return lib().cb200_mlkem_encaps(self._k, ek, stride, seeds, ct, ss, status, n)

    def _c_decaps(self, dk, stride, ct, ss, status, n):
        if self._round3:
            return lib().cb200_kyber_kem_decaps(self._k, dk, stride, ct, ss, n)
        return lib().cb200_mlkem_decaps(self._k, dk, stride, ct, ss, status, n)

    def _c_keygen(self, seeds, ek, dk, n):
        fn = lib().cb200_kyber_kem_keygen if self._round3 else lib().cb200_mlkem_keygen
        return fn(self._k, seeds, ek, dk, n)

    # ---- kem.Scheme size/identity methods (kyber.go:271-279) ----
    def Name(self) -> str:
        return self._name

    def PublicKeySize(self) -> int:
        return 384 * self._k + 32

    def PrivateKeySize(self) -> int:
        return 768 * self._k + 96

    def CiphertextSize(self) -> int:
        return {2: 768, 3: 1088, 4: 1568}[self._k]

    def SharedKeySize(self) -> int:
        return 32

    def SeedSize(self) -> int:
        return 64

    def EncapsulationSeedSize(self) -> int:
        return 32

    # ---- keys ----
    def UnmarshalBinaryPublicKey(self, buf: bytes) -> PublicKey:
        """kyber.go:390-396.  Length errors are reported here; the FIPS 203 modulus
        check (cpapke.go:45-55) is evaluated on the device when the key is first
        used and surfaces as ErrPubKey from Encapsulate*."""
        if len(buf) != self.PublicKeySize():
            raise ErrPubKeySize("kem: invalid public key size")
        return PublicKey(self, buf)

    # ---- encapsulation ----
    def EncapsulateDeterministically(self, pk: PublicKey, seed: bytes):
        """kyber.go:359-374 (batch of one)."""
        if not isinstance(pk, PublicKey) or pk._scheme is not self:
            raise ErrTypeMismatch("kem: type mismatch")
        if len(seed) != self.EncapsulationSeedSize():
            raise ErrSeedSize("kem: invalid seed size")
        ct, ss = self.EncapsulateBatch(pk, np.frombuffer(seed, dtype=np.uint8).reshape(1, 32))
        return ct[0].tobytes(), ss[0].tobytes()

    def EncapsulateBatch(self, pks, seeds, ct=None, ss=None, push=None):
        """Batched EncapsulateDeterministically.

        pks:   one PublicKey (shared by all ops), or an (n, PublicKeySize) uint8
               array / CUDA tensor of packed keys (one per op; A^T and H(ek) are
               rebuilt on the device for every op).
        seeds: (n, 32) uint8 array or CUDA tensor.
        Returns (ct, ss): (n, CiphertextSize), (n, 32) in the same kind of memory.
        Raises ErrPubKey if any key is not canonical.
        push:  (ct_ptr, ss_ptr) raw device addresses (CUDA tensors only): the rows are also copied there sub-batch by
               sub-batch while the batch computes (cb200_mlkem_encaps_push; the gather to rank 0 of circl_b200.shard).
        """
        k, eksz, ctsz = self._k, self.PublicKeySize(), self.CiphertextSize()
        torch_mode = _is_torch(seeds)
        if torch_mode:
            import torch
            n = seeds.shape[0]
            assert seeds.is_cuda and seeds.dtype == torch.uint8 and seeds.is_contiguous() and seeds.shape[1] == 32
            if isinstance(pks, PublicKey):
                ek = torch.frombuffer(bytearray(pks._packed), dtype=torch.uint8).to(seeds.device)
                stride = 0
            else:
                ek = pks
                assert ek.is_cuda and ek.is_contiguous() and tuple(ek.shape) == (n, eksz)
                stride = eksz
            ct = torch.empty((n, ctsz), dtype=torch.uint8, device=seeds.device) if ct is None else ct
            ss = torch.empty((n, 32), dtype=torch.uint8, device=seeds.device) if ss is None else ss
            status = torch.zeros((n,), dtype=torch.uint8, device=seeds.device)
            check(lib().cb200_set_stream(torch.cuda.current_stream().cuda_stream))
            if push is None:
                check(self._c_encaps(ek.data_ptr(), stride, seeds.data_ptr(), ct.data_ptr(), ss.data_ptr(),
                                     status.data_ptr(), n))
            else:
                check(lib().cb200_mlkem_encaps_push(self._k, ek.data_ptr(), stride, seeds.data_ptr(), ct.data_ptr(),
                                                    ss.data_ptr(), status.data_ptr(), n, push[0], push[1]))
            self._last_status = status  # read lazily: the call is asynchronous on the torch stream
            return ct, ss
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8)
        if seeds.ndim != 2 or seeds.shape[1] != 32:
            raise ErrSeedSize("kem: invalid seed size")
        n = seeds.shape[0]
        if isinstance(pks, PublicKey):
            ek = np.frombuffer(pks._packed, dtype=np.uint8)
            stride = 0
        else:
            ek = np.ascontiguousarray(pks, dtype=np.uint8)
            if ek.shape != (n, eksz):
                raise ErrPubKeySize("kem: invalid public key size")
            stride = eksz
        ct = np.empty((n, ctsz), dtype=np.uint8) if ct is None else ct
        ss = np.empty((n, 32), dtype=np.uint8) if ss is None else ss
        status = np.zeros((n,), dtype=np.uint8)
        try:
            check(self._c_encaps(ek.ctypes.data, stride, seeds.ctypes.data, ct.ctypes.data,
                                 ss.ctypes.data, status.ctypes.data, n))
        except Cb200Error as e:
            if e.code == -3:
                err = ErrPubKey("kem: invalid public key")
                err.status = status
                raise err from None
            raise
        return ct, ss

    def check_last_status(self):
        """Device-pointer mode: synchronise and raise ErrPubKey if any op of the last batch failed."""
        st = getattr(self, "_last_status", None)
        if st is not None and bool(st.any().item()):
            err = ErrPubKey("kem: invalid public key")
            err.status = st.cpu().numpy()
            raise err

    # ---- key generation ----
    def DeriveKeyPair(self, seed: bytes):
        """kyber.go:337-346: deterministic key pair from a 64-byte seed d || z (batch of one)."""
        if len(seed) != self.SeedSize():
            raise ValueError("kem: invalid seed size")  # the reference panics with kem.ErrSeedSize
        ek, dk = self.DeriveKeyPairBatch(np.frombuffer(seed, dtype=np.uint8).reshape(1, 64))
        return PublicKey(self, ek[0].tobytes()), PrivateKey(self, dk[0].tobytes())

    def GenerateKeyPair(self):
        """kyber.go:281-283: random seed from the OS, then DeriveKeyPair."""
        import os
        return self.DeriveKeyPair(os.urandom(self.SeedSize()))

    def DeriveKeyPairBatch(self, seeds):
        """seeds: (n, 64) uint8 array or CUDA tensor -> (ek (n, PublicKeySize), dk (n, PrivateKeySize)) packed keys."""
        k, eksz, dksz = self._k, self.PublicKeySize(), self.PrivateKeySize()
        if _is_torch(seeds):
            import torch
            n = seeds.shape[0]
            assert seeds.is_cuda and seeds.is_contiguous() and tuple(seeds.shape) == (n, 64)
            ek = torch.empty((n, eksz), dtype=torch.uint8, device=seeds.device)
            dk = torch.empty((n, dksz), dtype=torch.uint8, device=seeds.device)
            check(lib().cb200_set_stream(torch.cuda.current_stream().cuda_stream))
            check(self._c_keygen(seeds.data_ptr(), ek.data_ptr(), dk.data_ptr(), n))
            return ek, dk
        seeds = np.ascontiguousarray(seeds, dtype=np.uint8)
        if seeds.ndim != 2 or seeds.shape[1] != 64:
            raise ValueError("kem: invalid seed size")
        n = seeds.shape[0]
        ek = np.empty((n, eksz), dtype=np.uint8)
        dk = np.empty((n, dksz), dtype=np.uint8)
        check(self._c_keygen(seeds.ctypes.data, ek.ctypes.data, dk.ctypes.data, n))
        return ek, dk

    def UnmarshalBinaryPrivateKey(self, buf: bytes) -> PrivateKey:
        """kyber.go:398-407.  The H(ek) consistency check of PrivateKey.Unpack (kyber.go:226-228) is
        evaluated on the device at first use and surfaces as ErrPrivKey from Decapsulate*."""
        if len(buf) != self.PrivateKeySize():
            raise ErrPrivKeySize("kem: invalid private key size")
        return PrivateKey(self, buf)

    def Decapsulate(self, sk: PrivateKey, ct: bytes) -> bytes:
        """kyber.go:376-388 (batch of one)."""
        if not isinstance(sk, PrivateKey) or sk._scheme is not self:
            raise ErrTypeMismatch("kem: type mismatch")
        if len(ct) != self.CiphertextSize():
            raise ErrCiphertextSize("kem: invalid ciphertext size")
        return self.DecapsulateBatch(sk, np.frombuffer(ct, dtype=np.uint8).reshape(1, -1))[0].tobytes()

    def DecapsulateBatch(self, sks, cts, ss=None):
        """Batched Decapsulate.  sks: one PrivateKey (shared) or (n, PrivateKeySize) uint8 array / CUDA
        tensor; cts: (n, CiphertextSize).  Returns (n, 32) shared secrets (implicit rejection included)."""
        k, dksz, ctsz = self._k, self.PrivateKeySize(), self.CiphertextSize()
        if _is_torch(cts):
            import torch
            n = cts.shape[0]
            assert cts.is_cuda and cts.is_contiguous() and tuple(cts.shape) == (n, ctsz)
            assert not isinstance(sks, PrivateKey), "device path expects one dk per op"
            assert sks.is_cuda and sks.is_contiguous() and tuple(sks.shape) == (n, dksz)
            ss = torch.empty((n, 32), dtype=torch.uint8, device=cts.device) if ss is None else ss
            status = torch.zeros((n,), dtype=torch.uint8, device=cts.device)
            check(lib().cb200_set_stream(torch.cuda.current_stream().cuda_stream))
            check(self._c_decaps(sks.data_ptr(), dksz, cts.data_ptr(), ss.data_ptr(), status.data_ptr(), n))
            self._last_status = status
            return ss
        cts = np.ascontiguousarray(cts, dtype=np.uint8)
        if cts.ndim != 2 or cts.shape[1] != ctsz:
            raise ErrCiphertextSize("kem: invalid ciphertext size")
        n = cts.shape[0]
        if isinstance(sks, PrivateKey):
            dk = np.frombuffer(sks._packed, dtype=np.uint8)
            stride = 0
        else:
            dk = np.ascontiguousarray(sks, dtype=np.uint8)
            if dk.shape != (n, dksz):
                raise ErrPrivKeySize("kem: invalid private key size")
            stride = dksz
        ss = np.empty((n, 32), dtype=np.uint8) if ss is None else ss
        status = np.zeros((n,), dtype=np.uint8)
        try:
            check(self._c_decaps(dk.ctypes.data, stride, cts.ctypes.data, ss.ctypes.data,
                                 status.ctypes.data, n))
        except Cb200Error as e:
            if e.code == -5:
                err = ErrPrivKey("kem: invalid private key")
                err.status = status
                raise err from None
            raise
        return ss

_SCHEMES = {"ml-kem-512": Scheme("ML-KEM-512", 2), "ml-kem-768": Scheme("ML-KEM-768", 3),
            "ml-kem-1024": Scheme("ML-KEM-1024", 4),
            # round-3 Kyber (kem/kyber/kyber{512,768,1024}); same sizes, different FO wrapper
            "kyber512": Scheme("Kyber512", 2, True), "kyber768": Scheme("Kyber768", 3, True),
            "kyber1024": Scheme("Kyber1024", 4, True)}


def ByName(name: str):
    """kem/schemes/schemes.go:70 -- case-insensitive lookup; None if unknown."""
    return _SCHEMES.get(name.lower())


def All():
    return list(_SCHEMES.values())
