"""GPU parity: batched ML-KEM-768/1024 Encapsulate vs the NIST ACVP vectors and the oracle.

Reads like kem/mlkem/acvp_test.go:83-125 (UnmarshalBinaryPublicKey +
EncapsulateDeterministically against c/k) and kem/schemes/schemes_test.go.
"""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KS = {"ML-KEM-512": 2, "ML-KEM-768": 3, "ML-KEM-1024": 4}
Q = 3329


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def _h(tag, i, n):
    return hashlib.shake_256(bytes([tag]) + i.to_bytes(8, "little")).digest(n)


def key_pool(k, count):
    import oracle
    return [oracle.mlkem_keygen(k, _h(0, j, 64))[0] for j in range(count)]


@pytest.mark.parametrize("ps", list(KS))
def test_acvp_encaps_single_calls(cb, mlkem_acvp, ps):
    from circl_b200 import mlkem
    scheme = mlkem.ByName(ps)
    assert scheme is not None and scheme.Name() == ps
    for t in mlkem_acvp["encap"][ps][:5]:
        ek = scheme.UnmarshalBinaryPublicKey(bytes.fromhex(t["ek"]))
        ct, ss = scheme.EncapsulateDeterministically(ek, bytes.fromhex(t["m"]))
        assert ct.hex().upper() == t["c"].upper()
        assert ss.hex().upper() == t["k"].upper()


@pytest.mark.parametrize("ps", list(KS))
def test_acvp_encaps_one_batch(cb, mlkem_acvp, ps):
    from circl_b200 import mlkem
    scheme = mlkem.ByName(ps)
    tests = mlkem_acvp["encap"][ps]
    eks = np.stack([np.frombuffer(bytes.fromhex(t["ek"]), dtype=np.uint8) for t in tests])
    ms = np.stack([np.frombuffer(bytes.fromhex(t["m"]), dtype=np.uint8) for t in tests])
    ct, ss = scheme.EncapsulateBatch(eks, ms)
    for i, t in enumerate(tests):
        assert ct[i].tobytes().hex().upper() == t["c"].upper(), t["tcId"]
        assert ss[i].tobytes().hex().upper() == t["k"].upper(), t["tcId"]


@pytest.mark.parametrize("ps", list(KS))
@pytest.mark.parametrize("n", [1, 7, 33, 2000])
def test_batch_vs_oracle_per_op_keys(cb, ps, n):
    import oracle
    from circl_b200 import mlkem
    k = KS[ps]
    scheme = mlkem.ByName(ps)
    pool = key_pool(k, min(n, 16))
    eks = np.stack([np.frombuffer(pool[i % len(pool)], dtype=np.uint8) for i in range(n)])
    ms = np.stack([np.frombuffer(_h(1, i, 32), dtype=np.uint8) for i in range(n)])
    ct, ss = scheme.EncapsulateBatch(eks, ms)
    wct, wss, fails = oracle.mlkem_encaps_batch(k, eks, ms, nthreads=8)
    assert fails == 0
    assert np.array_equal(ct, wct) and np.array_equal(ss, wss)


@pytest.mark.parametrize("ps", list(KS))
def test_batch_vs_oracle_shared_key(cb, ps):
    import oracle
    from circl_b200 import mlkem
    k = KS[ps]
    scheme = mlkem.ByName(ps)
    ek = key_pool(k, 1)[0]
    n = 20000  # spans several L2-resident sub-batches
    ms = np.frombuffer(hashlib.shake_256(b"seeds").digest(32 * n), dtype=np.uint8).reshape(n, 32)
    ct, ss = scheme.EncapsulateBatch(scheme.UnmarshalBinaryPublicKey(ek), ms)
    wct, wss, fails = oracle.mlkem_encaps_batch(k, np.frombuffer(ek, dtype=np.uint8), ms, nthreads=8)
    assert fails == 0
    assert np.array_equal(ct, wct) and np.array_equal(ss, wss)


def test_non_canonical_key_is_err_pubkey(cb, mlkem_acvp):
    from circl_b200 import mlkem
    scheme = mlkem.ByName("ML-KEM-768")
    tests = mlkem_acvp["encap"]["ML-KEM-768"][:4]
    eks = np.stack([np.frombuffer(bytes.fromhex(t["ek"]), dtype=np.uint8) for t in tests]).copy()
    eks[2, 0], eks[2, 1] = 0xFF, eks[2, 1] | 0x0F  # first coefficient of op 2 = 4095 >= q
    ms = np.stack([np.frombuffer(bytes.fromhex(t["m"]), dtype=np.uint8) for t in tests])
    with pytest.raises(mlkem.ErrPubKey) as ei:
        scheme.EncapsulateBatch(eks, ms)
    assert ei.value.status.tolist() == [0, 0, 1, 0]
    with pytest.raises(mlkem.ErrPubKeySize):
        scheme.UnmarshalBinaryPublicKey(b"\x00" * 10)
    with pytest.raises(mlkem.ErrSeedSize):
        scheme.EncapsulateDeterministically(scheme.UnmarshalBinaryPublicKey(eks[0].tobytes()), b"\x00" * 31)


def test_device_pointers_large_batch_property(cb):
    """2^18 ops with device-resident buffers (per-op keys from a pool of 64):
    oracle-checked on a strided sample; every ciphertext must also decapsulate
    (oracle) to the same shared secret on that sample."""
    import torch
    import oracle
    from circl_b200 import mlkem
    k, n = 3, 1 << 18
    scheme = mlkem.ByName("ML-KEM-768")
    seeds64 = [_h(0, j, 64) for j in range(64)]
    keys = [oracle.mlkem_keygen(k, s) for s in seeds64]
    pool = torch.from_numpy(np.stack([np.frombuffer(ek, dtype=np.uint8) for ek, _ in keys])).cuda()
    idx = torch.arange(n, device="cuda") % 64
    eks = pool[idx].contiguous()
    g = torch.Generator(device="cuda").manual_seed(5)
    ms = torch.randint(0, 256, (n, 32), generator=g, device="cuda", dtype=torch.uint8)
    ct, ss = scheme.EncapsulateBatch(eks, ms)
    scheme.check_last_status()
    sample = list(range(0, n, 9973)) + [n - 1]
    ct_h, ss_h, ms_h = ct.cpu().numpy(), ss.cpu().numpy(), ms.cpu().numpy()
    for i in sample:
        ek, dk = keys[i % 64]
        wct, wss = oracle.mlkem_encaps(k, ek, ms_h[i].tobytes())
        assert ct_h[i].tobytes() == wct and ss_h[i].tobytes() == wss
        assert oracle.mlkem_decaps(k, dk, ct_h[i].tobytes()) == wss
    # checksum of checksums is stable across a second run (no races in the pipeline)
    ct2, ss2 = scheme.EncapsulateBatch(eks, ms)
    assert torch.equal(ct, ct2) and torch.equal(ss, ss2)


# ---------------------------------------------------------------- Decapsulate
@pytest.mark.parametrize("ps", list(KS))
def test_acvp_decaps(cb, mlkem_acvp, ps):
    # kem/mlkem/acvp_test.go:126-165 (VAL group: one dk, 10 ciphertexts incl. modified ones -> implicit rejection)
    from circl_b200 import mlkem
    scheme = mlkem.ByName(ps)
    g = mlkem_acvp["decap"][ps]
    sk = scheme.UnmarshalBinaryPrivateKey(bytes.fromhex(g["dk"]))
    cts = np.stack([np.frombuffer(bytes.fromhex(t["c"]), dtype=np.uint8) for t in g["tests"]])
    ss = scheme.DecapsulateBatch(sk, cts)
    for i, t in enumerate(g["tests"]):
        assert ss[i].tobytes().hex().upper() == t["k"].upper(), t["tcId"]
    assert scheme.Decapsulate(sk, cts[0].tobytes()).hex().upper() == g["tests"][0]["k"].upper()


@pytest.mark.parametrize("ps", list(KS))
def test_encaps_decaps_roundtrip_and_rejection(cb, ps):
    import oracle
    from circl_b200 import mlkem
    k = KS[ps]
    scheme = mlkem.ByName(ps)
    n = 3000
    keys = [oracle.mlkem_keygen(k, _h(0, j, 64)) for j in range(8)]
    eks = np.stack([np.frombuffer(keys[i % 8][0], dtype=np.uint8) for i in range(n)])
    dks = np.stack([np.frombuffer(keys[i % 8][1], dtype=np.uint8) for i in range(n)])
    ms = np.stack([np.frombuffer(_h(1, i, 32), dtype=np.uint8) for i in range(n)])
    ct, ss = scheme.EncapsulateBatch(eks, ms)
    assert np.array_equal(scheme.DecapsulateBatch(dks, ct), ss)          # kem round trip on the GPU alone
    bad = ct.copy()
    bad[::3, 5] ^= 0x40                                                   # corrupt every third ciphertext
    got = scheme.DecapsulateBatch(dks, bad)
    for i in list(range(0, 60)) + [n - 1]:
        assert got[i].tobytes() == oracle.mlkem_decaps(k, keys[i % 8][1], bad[i].tobytes()), i
    assert not np.array_equal(got[0], ss[0]) and np.array_equal(got[1], ss[1])
    # kem.ErrPrivKey: H(ek) stored in dk does not match
    broken = dks[:4].copy()
    broken[2, 384 * k + 384 * k + 32 + 1] ^= 1
    with pytest.raises(mlkem.ErrPrivKey) as ei:
        scheme.DecapsulateBatch(broken, ct[:4])
    assert ei.value.status.tolist() == [0, 0, 2, 0]


@pytest.mark.parametrize("ps", list(KS))
def test_decaps_with_unnormalised_secret_key_coefficients(cb, ps):
    """PrivateKey.Unpack normalises s-hat after the 12-bit unpack (cpapke.go:32-36), so a dk whose s-hat coefficients
    carry an extra q (still 12 bits) decapsulates like the canonical one.  The decrypt kernel feeds the packed words
    straight into its products (only residues matter): it must agree with the oracle, which follows the reference."""
    import oracle
    from circl_b200 import mlkem
    k = KS[ps]
    scheme = mlkem.ByName(ps)
    ek, dk = oracle.mlkem_keygen(k, _h(7, k, 64))
    dkb = bytearray(dk)
    changed = 0
    for i in range(0, 128 * k):  # pairs of coefficients, 3 bytes each
        b0, b1, b2 = dkb[3 * i], dkb[3 * i + 1], dkb[3 * i + 2]
        c0, c1 = b0 | ((b1 & 0xF) << 8), (b1 >> 4) | (b2 << 4)
        if c0 + Q < 4096 and i % 3 == 0:
            c0 += Q
            changed += 1
        if c1 + Q < 4096 and i % 5 == 0:
            c1 += Q
            changed += 1
        dkb[3 * i], dkb[3 * i + 1], dkb[3 * i + 2] = c0 & 0xFF, (c0 >> 8) | ((c1 & 0xF) << 4), c1 >> 4
    assert changed > 10
    n = 40
    ms = np.stack([np.frombuffer(_h(8, i, 32), dtype=np.uint8) for i in range(n)])
    eks = np.stack([np.frombuffer(ek, dtype=np.uint8)] * n)
    ct, ss = scheme.EncapsulateBatch(eks, ms)
    ct[::4, 9] ^= 0x11                                                    # some implicit rejections as well
    dks = np.stack([np.frombuffer(bytes(dkb), dtype=np.uint8)] * n)
    got = scheme.DecapsulateBatch(dks, ct)
    for i in range(n):
        assert got[i].tobytes() == oracle.mlkem_decaps(k, bytes(dkb), ct[i].tobytes()), i
    assert np.array_equal(got[1], ss[1])                                  # and it still is the key pair's secret


# ---------------------------------------------------------------- KeyGen
@pytest.mark.parametrize("ps", list(KS))
def test_acvp_keygen(cb, mlkem_acvp, ps):
    # kem/mlkem/acvp_test.go:41-82: seed = d || z -> ek, dk
    from circl_b200 import mlkem
    scheme = mlkem.ByName(ps)
    tests = mlkem_acvp["keygen"][ps]
    seeds = np.stack([np.frombuffer(bytes.fromhex(t["d"]) + bytes.fromhex(t["z"]), dtype=np.uint8) for t in tests])
    ek, dk = scheme.DeriveKeyPairBatch(seeds)
    for i, t in enumerate(tests):
        assert ek[i].tobytes().hex().upper() == t["ek"].upper(), t["tcId"]
        assert dk[i].tobytes().hex().upper() == t["dk"].upper(), t["tcId"]
    pk, sk = scheme.DeriveKeyPair(seeds[0].tobytes())
    assert pk.MarshalBinary() == ek[0].tobytes() and sk.MarshalBinary() == dk[0].tobytes()
    assert sk.Public().Equal(pk)


@pytest.mark.parametrize("ps", list(KS))
def test_keygen_encaps_decaps_all_on_gpu(cb, ps):
    """kem/schemes/schemes_test.go:53 style API round trip, 10000 key pairs, nothing but the GPU path."""
    import oracle
    from circl_b200 import mlkem
    k = KS[ps]
    scheme = mlkem.ByName(ps)
    n = 10000
    seeds = np.frombuffer(hashlib.shake_256(b"keygen" + ps.encode()).digest(64 * n), dtype=np.uint8).reshape(n, 64)
    ek, dk = scheme.DeriveKeyPairBatch(seeds)
    for i in range(0, n, 997):
        wek, wdk = oracle.mlkem_keygen(k, seeds[i].tobytes())
        assert ek[i].tobytes() == wek and dk[i].tobytes() == wdk
    ms = np.frombuffer(hashlib.shake_256(b"m").digest(32 * n), dtype=np.uint8).reshape(n, 32)
    ct, ss = scheme.EncapsulateBatch(ek, ms)
    assert np.array_equal(scheme.DecapsulateBatch(dk, ct), ss)
