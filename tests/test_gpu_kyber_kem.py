"""GPU parity: round-3 Kyber512/768/1024 KEM.

The reference pins these schemes by the SHA-256 of the PQCgenKAT transcript
(kem/kyber/kat_test.go:21-94); the first test rebuilds that transcript with every
key pair, ciphertext and shared secret computed on the GPU.  The others compare
against the oracle on seeded inputs, including implicit rejection.
"""
import hashlib

import numpy as np
import pytest

from nist_drbg import DRBG

pytestmark = pytest.mark.gpu

KS = {"Kyber512": 2, "Kyber768": 3, "Kyber1024": 4}


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def _h(tag, i, n):
    return hashlib.shake_256(bytes([tag]) + i.to_bytes(8, "little")).digest(n)


@pytest.mark.parametrize("name", list(KS))
def test_pqcgenkat_transcript_hash_on_gpu(cb, sampler_vectors, name):
    from circl_b200 import mlkem
    s = mlkem.ByName(name)
    assert s.Name() == name
    g = DRBG(bytes(range(48)))
    seeds, kseeds, eseeds = [], [], []
    for _ in range(100):
        seed = g.fill(48)
        g2 = DRBG(seed)
        seeds.append(seed)
        kseeds.append(g2.fill(32) + g2.fill(32))
        eseeds.append(g2.fill(32))
    ek, dk = s.DeriveKeyPairBatch(np.frombuffer(b"".join(kseeds), dtype=np.uint8).reshape(100, 64))
    ct, ss = s.EncapsulateBatch(ek, np.frombuffer(b"".join(eseeds), dtype=np.uint8).reshape(100, 32))
    ss2 = s.DecapsulateBatch(dk, ct)
    assert (ss == ss2).all()
    f = hashlib.sha256()
    f.update(("# %s\n\n" % name).encode())
    for i in range(100):
        f.update(("count = %d\nseed = %s\n" % (i, seeds[i].hex().upper())).encode())
        f.update(("pk = %s\nsk = %s\nct = %s\nss = %s\n\n" % (
            ek[i].tobytes().hex().upper(), dk[i].tobytes().hex().upper(), ct[i].tobytes().hex().upper(),
            ss[i].tobytes().hex().upper())).encode())
    assert f.hexdigest() == sampler_vectors["kat_sha256"][name]


@pytest.mark.parametrize("name", list(KS))
@pytest.mark.parametrize("n", [1, 33, 1500])
def test_batch_vs_oracle(cb, name, n):
    import oracle
    from circl_b200 import mlkem
    s, k = mlkem.ByName(name), KS[name]
    kseeds = np.frombuffer(b"".join(_h(0, j, 64) for j in range(n)), dtype=np.uint8).reshape(n, 64)
    eseeds = np.frombuffer(b"".join(_h(1, j, 32) for j in range(n)), dtype=np.uint8).reshape(n, 32)
    ek, dk = s.DeriveKeyPairBatch(kseeds)
    ct, ss = s.EncapsulateBatch(ek, eseeds)
    # implicit rejection: corrupt every third ciphertext
    bad = ct.copy()
    bad[::3, 5] ^= 0x40
    ss_good = s.DecapsulateBatch(dk, ct)
    ss_bad = s.DecapsulateBatch(dk, bad)
    assert (ss_good == ss).all()
    for i in list(range(min(n, 40))) + [n - 1]:
        oek, odk = oracle.kyber_kem_keygen(k, kseeds[i].tobytes())
        assert ek[i].tobytes() == oek and dk[i].tobytes() == odk, i
        oct_, oss = oracle.kyber_kem_encaps(k, oek, eseeds[i].tobytes())
        assert ct[i].tobytes() == oct_ and ss[i].tobytes() == oss, i
        assert ss_bad[i].tobytes() == oracle.kyber_kem_decaps(k, odk, bad[i].tobytes()), i
    assert (ss_bad[::3] != ss[::3]).any(axis=1).all()
    assert (ss_bad[1::3] == ss[1::3]).all()


@pytest.mark.parametrize("name", list(KS))
def test_single_calls_shared_key_and_lenient_parse(cb, name):
    import oracle
    from circl_b200 import mlkem
    s, k = mlkem.ByName(name), KS[name]
    pk, sk = s.DeriveKeyPair(_h(2, 0, 64))
    ct, ss = s.EncapsulateDeterministically(pk, _h(3, 0, 32))
    assert (ct, ss) == oracle.kyber_kem_encaps(k, pk.MarshalBinary(), _h(3, 0, 32))
    assert s.Decapsulate(sk, ct) == ss
    # one key shared by the whole batch (stride 0)
    seeds = np.frombuffer(b"".join(_h(4, j, 32) for j in range(50)), dtype=np.uint8).reshape(50, 32)
    cts, sss = s.EncapsulateBatch(pk, seeds)
    assert (s.DecapsulateBatch(sk, cts) == sss).all()
    assert (cts[7].tobytes(), sss[7].tobytes()) == oracle.kyber_kem_encaps(k, pk.MarshalBinary(), seeds[7].tobytes())
    # round-3 Unpack accepts unreduced coefficients (pke/kyber/internal/common/poly.go Unpack: no modulus check)
    raw = bytearray(pk.MarshalBinary())
    raw[0:3] = b"\xff\xff\xff"  # two 12-bit coefficients = 4095 > q
    pk2 = s.UnmarshalBinaryPublicKey(bytes(raw))
    assert s.EncapsulateDeterministically(pk2, _h(3, 1, 32)) == oracle.kyber_kem_encaps(k, bytes(raw), _h(3, 1, 32))


def test_device_pointers(cb):
    import torch
    import oracle
    from circl_b200 import mlkem
    s = mlkem.ByName("Kyber768")
    n = 20000
    g = torch.Generator(device="cpu").manual_seed(5)
    kseeds = torch.randint(0, 256, (n, 64), dtype=torch.uint8, generator=g).cuda()
    eseeds = torch.randint(0, 256, (n, 32), dtype=torch.uint8, generator=g).cuda()
    ek, dk = s.DeriveKeyPairBatch(kseeds)
    ct, ss = s.EncapsulateBatch(ek, eseeds)
    ss2 = s.DecapsulateBatch(dk, ct)
    torch.cuda.synchronize()
    assert torch.equal(ss, ss2)
    for i in (0, 8191, 8192, n - 1):
        oek, _ = oracle.kyber_kem_keygen(3, kseeds[i].cpu().numpy().tobytes())
        assert ek[i].cpu().numpy().tobytes() == oek
        assert (ct[i].cpu().numpy().tobytes(), ss[i].cpu().numpy().tobytes()) == \
            oracle.kyber_kem_encaps(3, oek, eseeds[i].cpu().numpy().tobytes())
