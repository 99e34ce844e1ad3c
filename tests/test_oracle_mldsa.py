"""Pins oracle/dilithium.c against every fixture the reference holds for the
Dilithium / ML-DSA-65 path (CPU only):
  * ACVP FIPS 204 keyGen / sigGen (deterministic + hedged, internal interface) / sigVer
        sign/mldsa/mldsa65/acvp_test.go:13-163
  * PQCgenKAT transcript SHA-256     sign/dilithium/kat_test.go:18-103
  * embedded sampler vectors         sign/mldsa/mldsa65/internal/sample_test.go:12, mode3/internal/params_test.go
  * algebraic self-checks            sign/internal/dilithium/ntt_test.go:25
"""
import hashlib

import numpy as np

import oracle
from nist_drbg import DRBG

Q = 8380417


def test_acvp_keygen(mldsa65_acvp):
    assert len(mldsa65_acvp["keygen"]) == 25
    for t in mldsa65_acvp["keygen"]:
        pk, sk = oracle.mldsa65_keygen(bytes.fromhex(t["seed"]))
        assert pk.hex().upper() == t["pk"].upper(), t["tcId"]
        assert sk.hex().upper() == t["sk"].upper(), t["tcId"]


def test_acvp_siggen(mldsa65_acvp):
    assert len(mldsa65_acvp["siggen"]) == 20
    for t in mldsa65_acvp["siggen"]:
        sig, attempts = oracle.mldsa65_sign(bytes.fromhex(t["sk"]), bytes.fromhex(t["message"]),
                                            rnd=bytes.fromhex(t["rnd"]), internal=True)
        assert sig.hex().upper() == t["signature"].upper(), t["tcId"]
        assert attempts >= 1


def test_acvp_sigver(mldsa65_acvp):
    g = mldsa65_acvp["sigver"]
    pk = bytes.fromhex(g["pk"])
    seen = set()
    for t in g["tests"]:
        got = oracle.mldsa65_verify(pk, bytes.fromhex(t["message"]), bytes.fromhex(t["signature"]), internal=True)
        assert got == t["testPassed"], t["tcId"]
        seen.add(got)
    assert seen == {True, False}


def test_pqcgenkat_hash(sampler_vectors):
    g = DRBG(bytes(range(48)))
    f = hashlib.sha256()
    f.update(b"# Dilithium3\n\n")
    for i in range(100):
        mlen = 33 * (i + 1)
        seed = g.fill(48)
        msg = g.fill(mlen)
        f.update(("count = %d\n" % i).encode())
        f.update(("seed = %s\n" % seed.hex().upper()).encode())
        f.update(("mlen = %d\n" % mlen).encode())
        f.update(("msg = %s\n" % msg.hex().upper()).encode())
        eseed = DRBG(seed).fill(32)
        pk, sk = oracle.mldsa65_keygen(eseed)
        f.update(("pk = %s\n" % pk.hex().upper()).encode())
        f.update(("sk = %s\n" % sk.hex().upper()).encode())
        f.update(("smlen = %d\n" % (mlen + 3309)).encode())
        sig, _ = oracle.mldsa65_sign(sk, msg)  # external interface, nil ctx, deterministic
        f.update(("sm = %s%s\n\n" % (sig.hex().upper(), msg.hex().upper())).encode())
        assert oracle.mldsa65_verify(pk, msg, sig)
    assert f.hexdigest() == sampler_vectors["kat_sha256"]["ML-DSA-65"]


def test_sampler_vectors(sampler_vectors):
    seed32, seed64 = bytes(range(32)), bytes(range(64))
    assert oracle.dil_derive_uniform(seed32, 30000).tolist() == sampler_vectors["dil_uniform_nonce30000"]
    # the mode3 vectors are compared after p.Normalize() (mode3/internal/params_test.go:38-39, :99-100)
    assert oracle.dil_poly_op(3, oracle.dil_derive_leqeta(seed64, 30000)).tolist() == sampler_vectors["dil_leqeta4_nonce30000"]
    assert oracle.dil_poly_op(3, oracle.dil_derive_legamma1(seed64, 30000)).tolist() == sampler_vectors["dil_legamma1_19_nonce30000"]


def test_ball_has_tau_signed_ones():
    # sample_test.go:105 TestDeriveUniformBall
    for i in range(50):
        p = oracle.dil_derive_ball(hashlib.shake_256(bytes([i])).digest(48))
        nz = p[p != 0]
        assert len(nz) == 49 and set(nz.tolist()) <= {1, Q - 1}


def test_zetas_spot_values():
    z, iz = oracle.dil_zetas()
    assert z[:3].tolist() == [4193792, 25847, 5771523]        # ntt.go:19
    assert iz[:3].tolist() == [6403635, 846154, 6979993]      # ntt.go:66
    assert iz[-1] == 4186625


def test_ntt_roundtrip():
    # ntt_test.go:25 TestNTT: InvNTT(NTT(p)) = R*p, bounds
    rng = np.random.default_rng(3)
    p = rng.integers(0, Q, size=(50, 256)).astype(np.uint32)
    ph = oracle.dil_ntt(p)
    assert ph.max() < 18 * Q
    back = oracle.dil_invntt(oracle.dil_poly_op(2, ph))
    assert back.max() < 2 * Q
    want = (p.astype(object) * (1 << 32)) % Q
    assert np.array_equal(oracle.dil_poly_op(3, back).astype(object), want)


def test_mulhat_is_negacyclic_product():
    rng = np.random.default_rng(4)
    a = rng.integers(0, Q, size=(4, 256)).astype(np.uint32)
    b = rng.integers(0, Q, size=(4, 256)).astype(np.uint32)
    ph = oracle.dil_mulhat(oracle.dil_ntt(a), oracle.dil_ntt(b))  # R^-1 * NTT(a)NTT(b)
    p = oracle.dil_poly_op(3, oracle.dil_invntt(oracle.dil_poly_op(2, ph)))  # InvNTT multiplies by R
    for i in range(4):
        full = np.convolve(a[i].astype(object), b[i].astype(object))
        full = np.concatenate([full, [0]])
        school = (full[:256] - full[256:]) % Q
        assert np.array_equal(p[i].astype(object), school)


# ---------------------------------------------------------------- ML-DSA-44 / ML-DSA-87 (run-time parameter set)
import pytest

MODES = {"ML-DSA-44": 44, "ML-DSA-87": 87}


@pytest.mark.parametrize("ps", list(MODES))
def test_other_modes_acvp(mldsa_other_acvp, ps):
    mode = MODES[ps]
    g = mldsa_other_acvp[ps]
    for t in g["keygen"]:
        pk, sk = oracle.mldsa_keygen(mode, bytes.fromhex(t["seed"]))
        assert pk.hex().upper() == t["pk"].upper() and sk.hex().upper() == t["sk"].upper(), t["tcId"]
    for t in g["siggen"]:
        sig, _ = oracle.mldsa_sign(mode, bytes.fromhex(t["sk"]), bytes.fromhex(t["message"]), rnd=bytes.fromhex(t["rnd"]),
                                   internal=True)
        assert sig.hex().upper() == t["signature"].upper(), t["tcId"]
    pk = bytes.fromhex(g["sigver"]["pk"])
    seen = set()
    for t in g["sigver"]["tests"]:
        got = oracle.mldsa_verify(mode, pk, bytes.fromhex(t["message"]), bytes.fromhex(t["signature"]), internal=True)
        assert got == t["testPassed"], t["tcId"]
        seen.add(got)
    assert seen == {True, False}


@pytest.mark.parametrize("ps", list(MODES))
def test_other_modes_pqcgenkat_hash(sampler_vectors, ps):
    # sign/dilithium/kat_test.go:43-100 (header names Dilithium2 / Dilithium5)
    mode = MODES[ps]
    g = DRBG(bytes(range(48)))
    f = hashlib.sha256()
    f.update(("# %s\n\n" % {44: "Dilithium2", 87: "Dilithium5"}[mode]).encode())
    _, _, sigsz = oracle.mldsa_sizes(mode)
    for i in range(100):
        mlen = 33 * (i + 1)
        seed = g.fill(48)
        msg = g.fill(mlen)
        f.update(("count = %d\nseed = %s\nmlen = %d\nmsg = %s\n" % (i, seed.hex().upper(), mlen, msg.hex().upper())).encode())
        pk, sk = oracle.mldsa_keygen(mode, DRBG(seed).fill(32))
        f.update(("pk = %s\nsk = %s\nsmlen = %d\n" % (pk.hex().upper(), sk.hex().upper(), mlen + sigsz)).encode())
        sig, _ = oracle.mldsa_sign(mode, sk, msg)
        f.update(("sm = %s%s\n\n" % (sig.hex().upper(), msg.hex().upper())).encode())
        assert oracle.mldsa_verify(mode, pk, msg, sig)
    assert f.hexdigest() == sampler_vectors["kat_sha256"][ps]


@pytest.mark.parametrize("name,mode", [("Dilithium2", 2), ("Dilithium3", 3), ("Dilithium5", 5)])
def test_round3_dilithium_pqcgenkat_hash(sampler_vectors, name, mode):
    # sign/dilithium/kat_test.go:25-27: round-3 parameter sets (NIST = false: tr and c~ of 32 bytes, no K/L in the key
    # seed hash, no rnd, raw message -- sign/dilithium/mode3/dilithium.go:54-77)
    g = DRBG(bytes(range(48)))
    f = hashlib.sha256()
    f.update(("# %s\n\n" % name).encode())
    _, _, sigsz = oracle.mldsa_sizes(mode)
    for i in range(100):
        mlen = 33 * (i + 1)
        seed = g.fill(48)
        msg = g.fill(mlen)
        f.update(("count = %d\nseed = %s\nmlen = %d\nmsg = %s\n" % (i, seed.hex().upper(), mlen, msg.hex().upper())).encode())
        pk, sk = oracle.mldsa_keygen(mode, DRBG(seed).fill(32))
        f.update(("pk = %s\nsk = %s\nsmlen = %d\n" % (pk.hex().upper(), sk.hex().upper(), mlen + sigsz)).encode())
        sig, _ = oracle.mldsa_sign(mode, sk, msg)
        f.update(("sm = %s%s\n\n" % (sig.hex().upper(), msg.hex().upper())).encode())
        assert oracle.mldsa_verify(mode, pk, msg, sig)
    assert f.hexdigest() == sampler_vectors["kat_sha256"][name]
