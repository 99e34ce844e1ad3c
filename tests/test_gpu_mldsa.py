"""GPU parity: batched ML-DSA-65 Sign vs the NIST ACVP vectors and the oracle (byte-exact).
Reads like sign/mldsa/mldsa65/acvp_test.go:81-123 and sign/schemes/schemes_test.go."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def _h(tag, i, n):
    return hashlib.shake_256(bytes([tag]) + i.to_bytes(8, "little")).digest(n)


def test_acvp_siggen_internal_interface(cb, mldsa65_acvp):
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    tests = mldsa65_acvp["siggen"]
    sks = np.stack([np.frombuffer(bytes.fromhex(t["sk"]), dtype=np.uint8) for t in tests])
    msgs = [bytes.fromhex(t["message"]) for t in tests]       # 246 .. 6877 bytes: multi-block absorb
    rnd = np.stack([np.frombuffer(bytes.fromhex(t["rnd"]), dtype=np.uint8) for t in tests])
    sig = scheme.SignBatch(sks, msgs, rnd=rnd, internal=True)
    for i, t in enumerate(tests):
        assert sig[i].tobytes().hex().upper() == t["signature"].upper(), t["tcId"]


@pytest.mark.parametrize("n", [1, 5, 300])
def test_batch_vs_oracle_per_op_keys(cb, n):
    import oracle
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    pool = [oracle.mldsa65_keygen(_h(2, j, 32)) for j in range(min(n, 8))]
    sks = np.stack([np.frombuffer(pool[i % len(pool)][1], dtype=np.uint8) for i in range(n)])
    msgs = [_h(3, i, 32 + (i % 7) * 50) for i in range(n)]
    sig, attempts = scheme.SignBatch(sks, msgs, return_attempts=True)
    want, want_attempts = oracle.mldsa65_sign_batch(sks, msgs, nthreads=8)
    assert np.array_equal(sig, want)
    assert attempts == want_attempts          # same rejection-loop trajectory, attempt for attempt
    for i in range(0, n, max(1, n // 10)):
        assert oracle.mldsa65_verify(pool[i % len(pool)][0], msgs[i], sig[i].tobytes())


def test_shared_key_context_and_hedged(cb):
    import oracle
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    pk, sk = oracle.mldsa65_keygen(_h(2, 99, 32))
    key = scheme.UnmarshalBinaryPrivateKey(sk)
    n = 200
    msgs = [_h(4, i, i % 90) for i in range(n)]               # includes the empty message
    ctx = b"circl-b200 test context"
    rnd = np.frombuffer(hashlib.shake_256(b"rnd").digest(32 * n), dtype=np.uint8).reshape(n, 32)
    sig = scheme.SignBatch(key, msgs, ctx=ctx, rnd=rnd)
    for i in range(0, n, 7):
        want, _ = oracle.mldsa65_sign(sk, msgs[i], ctx=ctx, rnd=rnd[i].tobytes())
        assert sig[i].tobytes() == want
        assert oracle.mldsa65_verify(pk, msgs[i], sig[i].tobytes(), ctx=ctx)
    # single-op sign.Scheme.Sign with SignatureOpts
    one = scheme.Sign(key, b"hello", mldsa.SignatureOpts(Context=b"ctx"))
    assert one == oracle.mldsa65_sign(sk, b"hello", ctx=b"ctx")[0]
    with pytest.raises(mldsa.ErrContextTooLong):
        scheme.SignBatch(key, [b"x"], ctx=b"\0" * 256)
    with pytest.raises(mldsa.ErrPrivKeySize):
        scheme.UnmarshalBinaryPrivateKey(b"\0" * 100)


def test_many_ops_property(cb):
    """2^13 signatures, shared key: all verify (oracle) on a sample; attempts ~ 5.1 per signature."""
    import oracle
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    pk, sk = oracle.mldsa65_keygen(_h(2, 7, 32))
    n = 1 << 13
    msgs = [_h(5, i, 32) for i in range(n)]
    sig, attempts = scheme.SignBatch(scheme.UnmarshalBinaryPrivateKey(sk), msgs, return_attempts=True)
    assert 3.5 < attempts / n < 7.0
    for i in range(0, n, 331):
        assert sig[i].tobytes() == oracle.mldsa65_sign(sk, msgs[i])[0]


# ---------------------------------------------------------------- Verify
def test_acvp_sigver(cb, mldsa65_acvp):
    # sign/mldsa/mldsa65/acvp_test.go:124-163: one pk, 15 signatures, expected pass / fail
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    g = mldsa65_acvp["sigver"]
    pk = scheme.UnmarshalBinaryPublicKey(bytes.fromhex(g["pk"]))
    msgs = [bytes.fromhex(t["message"]) for t in g["tests"]]
    sigs = np.stack([np.frombuffer(bytes.fromhex(t["signature"]), dtype=np.uint8) for t in g["tests"]])
    ok = scheme.VerifyBatch(pk, msgs, sigs, internal=True)
    assert ok.tolist() == [t["testPassed"] for t in g["tests"]]
    assert True in ok.tolist() and False in ok.tolist()


def test_sign_then_verify_on_gpu_and_tampering(cb):
    import oracle
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    n = 600
    pool = [oracle.mldsa65_keygen(_h(2, j, 32)) for j in range(6)]
    sks = np.stack([np.frombuffer(pool[i % 6][1], dtype=np.uint8) for i in range(n)])
    pks = np.stack([np.frombuffer(pool[i % 6][0], dtype=np.uint8) for i in range(n)])
    msgs = [_h(6, i, 1 + i % 200) for i in range(n)]
    ctx = b"ctx"
    sig = scheme.SignBatch(sks, msgs, ctx=ctx)
    assert scheme.VerifyBatch(pks, msgs, sig, ctx=ctx).all()
    assert not scheme.VerifyBatch(pks, msgs, sig, ctx=b"other").any()       # wrong context
    bad = sig.copy()
    bad[0, 3] ^= 1                 # c~
    bad[1, 48 + 7] ^= 0x10         # z
    bad[2, 3309 - 1] ^= 1          # hint switch-over point
    bad[3, 3309 - 20] = 200        # padding index not zero / ordering
    got = scheme.VerifyBatch(pks, msgs, bad, ctx=ctx)
    want = [oracle.mldsa65_verify(pool[i % 6][0], msgs[i], bad[i].tobytes(), ctx=ctx) for i in range(8)]
    assert got[:8].tolist() == want and not any(want[:4]) and all(want[4:])
    # shared public key + single-op API
    pk0 = scheme.UnmarshalBinaryPublicKey(pool[0][0])
    idx = list(range(0, n, 6))
    assert scheme.VerifyBatch(pk0, [msgs[i] for i in idx], sig[idx], ctx=ctx).all()
    assert scheme.Verify(pk0, msgs[0], sig[0].tobytes(), mldsa.SignatureOpts(Context=ctx))
    assert not scheme.Verify(pk0, msgs[0], sig[0].tobytes()[:-1], mldsa.SignatureOpts(Context=ctx))


# ---------------------------------------------------------------- KeyGen
def test_acvp_keygen(cb, mldsa65_acvp):
    # sign/mldsa/mldsa65/acvp_test.go:47-80: seed -> pk, sk
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    tests = mldsa65_acvp["keygen"]
    seeds = np.stack([np.frombuffer(bytes.fromhex(t["seed"]), dtype=np.uint8) for t in tests])
    pk, sk = scheme.DeriveKeyBatch(seeds)
    for i, t in enumerate(tests):
        assert pk[i].tobytes().hex().upper() == t["pk"].upper(), t["tcId"]
        assert sk[i].tobytes().hex().upper() == t["sk"].upper(), t["tcId"]


def test_keygen_sign_verify_all_on_gpu(cb):
    """sign/schemes/schemes_test.go:17 style API round trip with nothing but the GPU path."""
    import oracle
    from circl_b200 import mldsa
    scheme = mldsa.ByName("ML-DSA-65")
    n = 2000
    seeds = np.frombuffer(hashlib.shake_256(b"dsa-keygen").digest(32 * n), dtype=np.uint8).reshape(n, 32)
    pk, sk = scheme.DeriveKeyBatch(seeds)
    for i in range(0, n, 397):
        wpk, wsk = oracle.mldsa65_keygen(seeds[i].tobytes())
        assert pk[i].tobytes() == wpk and sk[i].tobytes() == wsk
    msgs = [_h(7, i, 40) for i in range(n)]
    sig = scheme.SignBatch(sk, msgs)
    assert scheme.VerifyBatch(pk, msgs, sig).all()
    assert not scheme.VerifyBatch(pk, msgs[1:] + msgs[:1], sig).any()
    p1, s1 = scheme.DeriveKey(seeds[0].tobytes())
    assert p1.MarshalBinary() == pk[0].tobytes() and scheme.Verify(p1, b"m", scheme.Sign(s1, b"m"))


# ---------------------------------------------------------------- ML-DSA-44 / ML-DSA-87
OTHER = {"ML-DSA-44": 44, "ML-DSA-87": 87}


@pytest.mark.parametrize("ps", list(OTHER))
def test_other_modes_acvp(cb, mldsa_other_acvp, ps):
    from circl_b200 import mldsa
    scheme = mldsa.ByName(ps)
    g = mldsa_other_acvp[ps]
    seeds = np.stack([np.frombuffer(bytes.fromhex(t["seed"]), dtype=np.uint8) for t in g["keygen"]])
    pk, sk = scheme.DeriveKeyBatch(seeds)
    for i, t in enumerate(g["keygen"]):
        assert pk[i].tobytes().hex().upper() == t["pk"].upper() and sk[i].tobytes().hex().upper() == t["sk"].upper()
    tests = g["siggen"]
    sks = np.stack([np.frombuffer(bytes.fromhex(t["sk"]), dtype=np.uint8) for t in tests])
    rnd = np.stack([np.frombuffer(bytes.fromhex(t["rnd"]), dtype=np.uint8) for t in tests])
    sig = scheme.SignBatch(sks, [bytes.fromhex(t["message"]) for t in tests], rnd=rnd, internal=True)
    for i, t in enumerate(tests):
        assert sig[i].tobytes().hex().upper() == t["signature"].upper(), t["tcId"]
    v = g["sigver"]
    ok = scheme.VerifyBatch(scheme.UnmarshalBinaryPublicKey(bytes.fromhex(v["pk"])),
                            [bytes.fromhex(t["message"]) for t in v["tests"]],
                            np.stack([np.frombuffer(bytes.fromhex(t["signature"]), dtype=np.uint8) for t in v["tests"]]),
                            internal=True)
    assert ok.tolist() == [t["testPassed"] for t in v["tests"]]


@pytest.mark.parametrize("ps", list(OTHER))
def test_other_modes_vs_oracle_and_roundtrip(cb, ps):
    import oracle
    from circl_b200 import mldsa
    mode = OTHER[ps]
    scheme = mldsa.ByName(ps)
    n = 500
    seeds = np.frombuffer(hashlib.shake_256(ps.encode()).digest(32 * n), dtype=np.uint8).reshape(n, 32)
    pk, sk = scheme.DeriveKeyBatch(seeds)
    for i in range(0, n, 101):
        wpk, wsk = oracle.mldsa_keygen(mode, seeds[i].tobytes())
        assert pk[i].tobytes() == wpk and sk[i].tobytes() == wsk
    msgs = [_h(8, i, 10 + i % 120) for i in range(n)]
    sig, attempts = scheme.SignBatch(sk, msgs, ctx=b"c", return_attempts=True)
    for i in range(0, n, 37):
        want, _ = oracle.mldsa_sign(mode, sk[i].tobytes(), msgs[i], ctx=b"c")
        assert sig[i].tobytes() == want
    want_all, want_attempts = oracle.mldsa_sign_batch(mode, sk[:64], msgs[:64], nthreads=8)
    sig64, att64 = scheme.SignBatch(sk[:64], msgs[:64], return_attempts=True)
    assert np.array_equal(sig64, want_all) and att64 == want_attempts
    assert scheme.VerifyBatch(pk, msgs, sig, ctx=b"c").all()
    bad = sig.copy()
    bad[:, 70] ^= 2
    got = scheme.VerifyBatch(pk, msgs, bad, ctx=b"c")
    assert got[:20].tolist() == [oracle.mldsa_verify(mode, pk[i].tobytes(), msgs[i], bad[i].tobytes(), ctx=b"c") for i in range(20)]
    assert not got.any()


ROUND3 = {"Dilithium2": 2, "Dilithium3": 3, "Dilithium5": 5}


@pytest.mark.parametrize("ps", list(ROUND3))
def test_round3_dilithium_pqcsignkat_transcript_on_gpu(cb, sampler_vectors, ps):
    # sign/dilithium/kat_test.go:18-100 with every key pair and signature computed by the CUDA path in three batches
    from nist_drbg import DRBG
    from circl_b200 import mldsa
    scheme = mldsa.ByName(ps)
    assert scheme.Name() == ps and not scheme.SupportsContext()
    g = DRBG(bytes(range(48)))
    seeds, msgs, eseeds = [], [], []
    for i in range(100):
        seeds.append(g.fill(48))
        msgs.append(g.fill(33 * (i + 1)))
        eseeds.append(DRBG(seeds[-1]).fill(32))
    pk, sk = scheme.DeriveKeyBatch(np.frombuffer(b"".join(eseeds), dtype=np.uint8).reshape(100, 32))
    sig = scheme.SignBatch(sk, msgs)
    assert scheme.VerifyBatch(pk, msgs, sig).all()
    f = hashlib.sha256()
    f.update(("# %s\n\n" % ps).encode())
    for i in range(100):
        mlen = len(msgs[i])
        f.update(("count = %d\nseed = %s\nmlen = %d\nmsg = %s\n" % (i, seeds[i].hex().upper(), mlen, msgs[i].hex().upper())).encode())
        f.update(("pk = %s\nsk = %s\nsmlen = %d\n" % (pk[i].tobytes().hex().upper(), sk[i].tobytes().hex().upper(),
                                                     mlen + scheme.SignatureSize())).encode())
        f.update(("sm = %s%s\n\n" % (sig[i].tobytes().hex().upper(), msgs[i].hex().upper())).encode())
    assert f.hexdigest() == sampler_vectors["kat_sha256"][ps]


@pytest.mark.parametrize("ps", list(ROUND3))
def test_round3_dilithium_vs_oracle(cb, ps):
    import oracle
    from circl_b200 import mldsa
    mode, scheme = ROUND3[ps], mldsa.ByName(ps)
    n = 300
    seeds = np.frombuffer(hashlib.shake_256(ps.encode()).digest(32 * n), dtype=np.uint8).reshape(n, 32)
    pk, sk = scheme.DeriveKeyBatch(seeds)
    msgs = [_h(9, i, 1 + i % 200) for i in range(n)]
    sig, attempts = scheme.SignBatch(sk, msgs, return_attempts=True)
    want, want_attempts = oracle.mldsa_sign_batch(mode, sk, msgs, nthreads=8)
    assert np.array_equal(sig, want) and attempts == want_attempts
    for i in range(0, n, 59):
        assert (pk[i].tobytes(), sk[i].tobytes()) == oracle.mldsa_keygen(mode, seeds[i].tobytes())
    # one shared key, single calls, tampering
    p0, s0 = scheme.DeriveKey(seeds[0].tobytes())
    one = scheme.Sign(s0, b"round three")
    assert one == oracle.mldsa_sign(mode, s0.MarshalBinary(), b"round three")[0]
    assert scheme.Verify(p0, b"round three", one) and not scheme.Verify(p0, b"round thre3", one)
    bad = sig.copy()
    bad[:, 40] ^= 1
    got = scheme.VerifyBatch(pk, msgs, bad)
    assert got[:16].tolist() == [oracle.mldsa_verify(mode, pk[i].tobytes(), msgs[i], bad[i].tobytes()) for i in range(16)]
    with pytest.raises(mldsa.ErrContextNotSupported):
        scheme.Sign(s0, b"m", mldsa.SignatureOpts(Context=b"ctx"))
