"""Pins the AVX2 arm of the oracle (oracle/kyber_avx2.c: the reference's amd64 fast path -- f1600x4AVX2, nttAVX2,
invNttAVX2, mulHatAVX2, PolyDeriveUniformX4 -- restated with intrinsics) against the generic restatement and against the
reference's own vectors.  CPU only.  Mirrors simd/keccakf1600/f1600x_test.go:21-57 (X4 equals four scalar permutations)
and pke/kyber/internal/common/{ntt,poly}_test.go (accelerated leaf == generic leaf on random inputs)."""
import ctypes as C

import numpy as np
import pytest

import oracle

Q = 3329


def _call(name, *arrays):
    getattr(oracle.lib(), name)(*[a.ctypes.data_as(C.c_void_p) for a in arrays])


def test_keccak_x4_equals_four_scalar_permutations(sampler_vectors):
    rng = np.random.default_rng(4)
    st = rng.integers(0, 1 << 63, size=(4, 25), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(4, 25), dtype=np.uint64)
    st[2] = 0  # one instance is the zero state of f1600x_test.go:9-19
    got = oracle.keccak_f1600_x4(st)
    for j in range(4):
        assert got[j].tolist() == oracle.keccak_f1600([int(x) for x in st[j]])
    assert got[2].tolist() == sampler_vectors["keccak_f1600_of_zero"]


def test_ntt_and_mulhat_equal_generic_coefficient_for_coefficient():
    rng = np.random.default_rng(5)
    p = rng.integers(-Q, Q, size=(200, 256), dtype=np.int64).astype(np.int16)
    for row in p:
        a = row.copy()
        _call("orc_kyber_ntt_avx2", a)
        assert np.array_equal(a, oracle.kyber_ntt(row))
    wild = rng.integers(-32768, 32768, size=(50, 256), dtype=np.int64).astype(np.int16)  # int16 wrap-around included
    for i in range(0, 50, 2):
        out = np.empty(256, dtype=np.int16)
        _call("orc_kyber_mulhat_avx2", out, np.ascontiguousarray(wild[i]), np.ascontiguousarray(wild[i + 1]))
        assert np.array_equal(out, oracle.kyber_mulhat(wild[i], wild[i + 1]))


def test_invntt_equals_generic_modulo_q():
    rng = np.random.default_rng(6)
    p = rng.integers(-Q, Q + 1, size=(200, 256), dtype=np.int64).astype(np.int16)
    p[0, :] = Q   # extreme inputs of the stated bound |x| <= q (ntt.go:145-150)
    p[1, :] = -Q
    for row in p:
        a = row.copy()
        _call("orc_kyber_invntt_avx2", a)
        assert int(np.abs(a).max()) < Q  # the output bound the reference states
        assert np.array_equal(oracle.kyber_normalize(a), oracle.kyber_normalize(oracle.kyber_invntt(row)))


@pytest.mark.parametrize("k", [2, 3, 4])
def test_encaps_bytes_equal_generic_arm(k):
    rng = np.random.default_rng(k)
    eks = np.stack([np.frombuffer(oracle.mlkem_keygen(k, bytes(rng.integers(0, 256, 64, dtype=np.uint8)))[0], dtype=np.uint8)
                    for _ in range(16)])
    n = 600  # ~1 % of the SHAKE128 streams need a fourth block: several X4 groups with uneven stream lengths
    e = np.ascontiguousarray(eks[np.arange(n) % 16])
    m = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    ct1, ss1, f1 = oracle.mlkem_encaps_batch(k, e, m, nthreads=2)
    ct2, ss2, f2 = oracle.mlkem_encaps_batch_avx2(k, e, m, nthreads=3)
    assert f1 == f2 == 0 and np.array_equal(ct1, ct2) and np.array_equal(ss1, ss2)
    e[5, :2] = 0xFF  # non-canonical key: kem.ErrPubKey in both arms
    assert oracle.mlkem_encaps_batch_avx2(k, e, m, nthreads=1)[2] == 1


def test_acvp_encapsulation_vectors_through_the_avx2_arm(mlkem_acvp):
    for ps, k in (("ML-KEM-512", 2), ("ML-KEM-768", 3), ("ML-KEM-1024", 4)):
        tests = mlkem_acvp["encap"][ps]
        ek = np.stack([np.frombuffer(bytes.fromhex(t["ek"]), dtype=np.uint8) for t in tests])
        m = np.stack([np.frombuffer(bytes.fromhex(t["m"]), dtype=np.uint8) for t in tests])
        ct, ss, fails = oracle.mlkem_encaps_batch_avx2(k, ek, m, nthreads=1)
        assert fails == 0
        for i, t in enumerate(tests):
            assert ct[i].tobytes().hex().upper() == t["c"].upper() and ss[i].tobytes().hex().upper() == t["k"].upper()
