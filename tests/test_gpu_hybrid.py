"""GPU parity for the wire-side callers: X25519 against the reference's own vectors
(dh/x25519/key_test.go), X-Wing against the draft vectors hash (kem/xwing/xwing_test.go:40-83), the kem/hybrid
schemes against the oracle, and the low-order-point behaviour of kem/hybrid/xkem_test.go."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden
from test_oracle_hybrid import xwing_transcript

pytestmark = pytest.mark.gpu

LOW = bytes.fromhex("e0eb7a7c3b41b8ae1656e3faf19fc46ada098deb9c32b1fd866205165f49b800")


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


@pytest.fixture(scope="module")
def xv():
    return load_golden("x25519_vectors.json.gz")


def _rows(items):
    return np.frombuffer(b"".join(items), dtype=np.uint8).reshape(len(items), -1)


def _h(tag, i, n):
    return hashlib.shake_256(bytes([tag]) + i.to_bytes(8, "little")).digest(n)


def test_x25519_rfc7748_and_wycheproof_one_batch(cb, xv):
    from circl_b200 import hybrid
    vec = [(v["scalar"], v["input"], v["output"], True) for v in xv["rfc7748_kat"]]
    vec += [(v["private"], v["public"], v["shared"], v["result"] != "acceptable") for v in xv["wycheproof"]]
    out, ok = hybrid.x25519_shared(_rows([bytes.fromhex(v[0]) for v in vec]), _rows([bytes.fromhex(v[1]) for v in vec]))
    for i, v in enumerate(vec):
        assert out[i].tobytes().hex() == v[2], i
        assert ok[i] or not v[3], i  # key_test.go:136-138: ok may only be false for "acceptable" vectors


def test_x25519_iterated_and_base_point(cb, xv):
    import oracle
    from circl_b200 import hybrid
    for v in xv["rfc7748_times"]:  # key_test.go:53-86
        u = k = bytes([9] + [0] * 31)
        for _ in range(v["times"]):
            r = hybrid.x25519_shared(_rows([k]), _rows([u]))[0][0].tobytes()
            u, k = k, r
        assert k.hex() == v["key"]
    n = 3000
    ks = _rows([_h(1, i, 32) for i in range(n)])
    pub = hybrid.x25519_keygen(ks)
    base = np.zeros((n, 32), dtype=np.uint8)
    base[:, 0] = 9
    shared, ok = hybrid.x25519_shared(ks, base)  # key_test.go:88-101
    assert ok.all() and (pub == shared).all()
    for i in range(0, n, 271):
        assert pub[i].tobytes() == oracle.x25519(ks[i].tobytes())[0]
    # second party: shared secrets agree and equal the oracle's
    ks2 = _rows([_h(2, i, 32) for i in range(n)])
    s12, _ = hybrid.x25519_shared(ks, hybrid.x25519_keygen(ks2))
    s21, _ = hybrid.x25519_shared(ks2, pub)
    assert (s12 == s21).all()
    assert s12[17].tobytes() == oracle.x25519(ks[17].tobytes(), oracle.x25519(ks2[17].tobytes())[0])[0]


def test_xwing_draft_vectors_on_gpu(cb, xv):
    from circl_b200 import hybrid
    s = hybrid.ByName("X-Wing")
    assert (s.PublicKeySize(), s.PrivateKeySize(), s.CiphertextSize(), s.EncapsulationSeedSize()) == (1216, 32, 1120, 64)

    def derive(seed):
        return s.DeriveKeyPair(seed)[0].MarshalBinary()

    def encaps(pk, eseed):
        return s.EncapsulateDeterministically(s.UnmarshalBinaryPublicKey(pk), eseed)

    def decaps(sk, ct):
        return s.Decapsulate(s.UnmarshalBinaryPrivateKey(sk), ct)

    assert xwing_transcript(derive, encaps, decaps) == xv["xwing_vectors_shake128"]


def test_xwing_batch_vs_oracle(cb):
    import oracle
    from circl_b200 import hybrid, mlkem
    s = hybrid.ByName("x-wing")
    n = 700
    seeds = _rows([_h(3, i, 32) for i in range(n)])
    eseeds = _rows([_h(4, i, 64) for i in range(n)])
    pk, sk = s.DeriveKeyPairBatch(seeds)
    ct, ss = s.EncapsulateBatch(pk, eseeds)
    assert (s.DecapsulateBatch(sk, ct) == ss).all()
    bad = ct.copy()
    bad[::2, 100] ^= 1          # ML-KEM half: implicit rejection
    bad[1::2, 1100] ^= 1        # X25519 half
    ss_bad = s.DecapsulateBatch(sk, bad)
    assert (ss_bad != ss).any(axis=1).all()
    for i in list(range(0, n, 53)) + [n - 1]:
        opk = oracle.xwing_keygen(seeds[i].tobytes())
        assert pk[i].tobytes() == opk
        assert (ct[i].tobytes(), ss[i].tobytes()) == oracle.xwing_encaps(opk, eseeds[i].tobytes())
        assert ss_bad[i].tobytes() == oracle.xwing_decaps(seeds[i].tobytes(), bad[i].tobytes())
    # one shared key for the whole batch; a non-canonical ML-KEM half is kem.ErrPubKey (xwing.go:173-177)
    p0, s0 = s.DeriveKeyPair(seeds[0].tobytes())
    c1, k1 = s.EncapsulateBatch(p0, eseeds[:40])
    assert (s.DecapsulateBatch(s0, c1) == k1).all()
    raw = bytearray(p0.MarshalBinary())
    raw[0:2] = b"\xff\x0f"
    with pytest.raises(mlkem.ErrPubKey):
        s.EncapsulateDeterministically(s.UnmarshalBinaryPublicKey(bytes(raw)), eseeds[0].tobytes())


@pytest.mark.parametrize("name", ["X25519MLKEM768", "Kyber768-X25519", "Kyber512-X25519"])
def test_hybrid_vs_oracle_and_low_order_points(cb, name):
    import oracle
    from circl_b200 import hybrid, mlkem
    s = hybrid.ByName(name)
    assert s.Name() == name
    assert (s.PublicKeySize(), s.PrivateKeySize(), s.CiphertextSize()) == oracle.hybrid_sizes(name)
    n = 400
    seeds = _rows([_h(5, i, 64) for i in range(n)])
    eseeds = _rows([_h(6, i, 32) for i in range(n)])
    pk, sk = s.DeriveKeyPairBatch(seeds)
    ct, ss = s.EncapsulateBatch(pk, eseeds)
    assert (s.DecapsulateBatch(sk, ct) == ss).all()
    for i in list(range(0, n, 37)) + [n - 1]:
        assert (pk[i].tobytes(), sk[i].tobytes()) == oracle.hybrid_keygen(name, seeds[i].tobytes())
        assert (ct[i].tobytes(), ss[i].tobytes(), 0) == oracle.hybrid_encaps(name, pk[i].tobytes(), eseeds[i].tobytes())
    # single calls with key objects; shared key batches
    p0, s0 = s.DeriveKeyPair(seeds[0].tobytes())
    c0, k0 = s.EncapsulateDeterministically(p0, eseeds[0].tobytes())
    assert (c0, k0) == (ct[0].tobytes(), ss[0].tobytes()) and s.Decapsulate(s0, c0) == k0
    cs, ks = s.EncapsulateBatch(p0, eseeds[:33])
    assert (s.DecapsulateBatch(s0, cs) == ks).all()
    # kem/hybrid/xkem_test.go:19-68: a small-order X25519 share is kem.ErrPubKey on both sides, per op
    x_first = name != "X25519MLKEM768"
    sl = slice(0, 32) if x_first else slice(-32, None)
    bad_pk = pk[:8].copy()
    bad_pk[3, sl] = np.frombuffer(LOW, dtype=np.uint8)
    with pytest.raises(mlkem.ErrPubKey) as ei:
        s.EncapsulateBatch(bad_pk, eseeds[:8])
    assert ei.value.status.tolist() == [0, 0, 0, 1, 0, 0, 0, 0]
    bad_ct = ct[:8].copy()
    bad_ct[5, sl] = np.frombuffer(LOW, dtype=np.uint8)
    with pytest.raises(mlkem.ErrPubKey) as ei:
        s.DecapsulateBatch(sk[:8], bad_ct)
    assert (ei.value.status != 0).tolist() == [False] * 5 + [True] + [False] * 2
    assert oracle.hybrid_decaps(name, sk[5].tobytes(), bad_ct[5].tobytes())[1] == 1


def test_device_pointers(cb):
    import torch
    import oracle
    from circl_b200 import hybrid
    n = 5000
    g = torch.Generator(device="cpu").manual_seed(11)
    xs = hybrid.ByName("X-Wing")
    seeds = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g).cuda()
    eseeds = torch.randint(0, 256, (n, 64), dtype=torch.uint8, generator=g).cuda()
    pk, sk = xs.DeriveKeyPairBatch(seeds)
    ct, ss = xs.EncapsulateBatch(pk, eseeds)
    ss2 = xs.DecapsulateBatch(sk, ct)
    torch.cuda.synchronize()
    assert torch.equal(ss, ss2)
    i = 4321
    assert (ct[i].cpu().numpy().tobytes(), ss[i].cpu().numpy().tobytes()) == \
        oracle.xwing_encaps(oracle.xwing_keygen(seeds[i].cpu().numpy().tobytes()), eseeds[i].cpu().numpy().tobytes())
    hs = hybrid.ByName("X25519MLKEM768")
    hseeds = torch.randint(0, 256, (n, 64), dtype=torch.uint8, generator=g).cuda()
    hpk, hsk = hs.DeriveKeyPairBatch(hseeds)
    hct, hss = hs.EncapsulateBatch(hpk, seeds)
    hss2 = hs.DecapsulateBatch(hsk, hct)
    torch.cuda.synchronize()
    assert torch.equal(hss, hss2)
    assert (hct[7].cpu().numpy().tobytes(), hss[7].cpu().numpy().tobytes(), 0) == \
        oracle.hybrid_encaps("X25519MLKEM768", hpk[7].cpu().numpy().tobytes(), seeds[7].cpu().numpy().tobytes())
