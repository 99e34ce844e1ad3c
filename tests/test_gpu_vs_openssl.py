"""GPU path against an implementation that shares nothing with the reference or this repository: OpenSSL's ML-KEM,
ML-DSA and X25519 through the `cryptography` package of the image.  Equal keys from equal seeds; each side
decapsulates / verifies what the other produced."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ossl_mlkem = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.mlkem")
ossl_mldsa = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.mldsa")
ossl_x25519 = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.x25519")


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def _rows(tag, n, width):
    return np.frombuffer(b"".join(hashlib.shake_256(bytes([tag]) + i.to_bytes(4, "little")).digest(width) for i in range(n)),
                         dtype=np.uint8).reshape(n, width)


@pytest.mark.parametrize("name,priv,pub", [("ML-KEM-768", "MLKEM768PrivateKey", "MLKEM768PublicKey"),
                                           ("ML-KEM-1024", "MLKEM1024PrivateKey", "MLKEM1024PublicKey")])
def test_mlkem(cb, name, priv, pub):
    from circl_b200 import mlkem
    s, n = mlkem.ByName(name), 24
    seeds, ms = _rows(1, n, 64), _rows(2, n, 32)
    ek, dk = s.DeriveKeyPairBatch(seeds)
    ct, ss = s.EncapsulateBatch(ek, ms)
    theirs = [getattr(ossl_mlkem, priv).from_seed_bytes(seeds[i].tobytes()) for i in range(n)]
    their_ct, their_ss = [], []
    for i in range(n):
        assert theirs[i].public_key().public_bytes_raw() == ek[i].tobytes()
        assert theirs[i].decapsulate(ct[i].tobytes()) == ss[i].tobytes()
        a, b = getattr(ossl_mlkem, pub).from_public_bytes(ek[i].tobytes()).encapsulate()
        k_, c_ = (a, b) if len(a) == 32 else (b, a)
        their_ct.append(c_)
        their_ss.append(k_)
    ours = s.DecapsulateBatch(dk, np.frombuffer(b"".join(their_ct), dtype=np.uint8).reshape(n, -1))
    assert [ours[i].tobytes() for i in range(n)] == their_ss


@pytest.mark.parametrize("name,priv", [("ML-DSA-44", "MLDSA44PrivateKey"), ("ML-DSA-65", "MLDSA65PrivateKey"),
                                       ("ML-DSA-87", "MLDSA87PrivateKey")])
def test_mldsa(cb, name, priv):
    from circl_b200 import mldsa
    s, n = mldsa.ByName(name), 16
    seeds = _rows(3, n, 32)
    pk, sk = s.DeriveKeyBatch(seeds)
    msgs = [hashlib.shake_256(b"msg%d" % i).digest(5 + 17 * i) for i in range(n)]
    sig = s.SignBatch(sk, msgs, ctx=b"tls13")
    theirs = [getattr(ossl_mldsa, priv).from_seed_bytes(seeds[i].tobytes()) for i in range(n)]
    their_sigs = []
    for i in range(n):
        assert theirs[i].public_key().public_bytes_raw() == pk[i].tobytes()
        theirs[i].public_key().verify(sig[i].tobytes(), msgs[i], b"tls13")  # raises on an invalid signature
        their_sigs.append(theirs[i].sign(msgs[i], b"tls13"))
    ok = s.VerifyBatch(pk, msgs, np.frombuffer(b"".join(their_sigs), dtype=np.uint8).reshape(n, -1), ctx=b"tls13")
    assert ok.all()
    assert not s.VerifyBatch(pk, msgs, np.frombuffer(b"".join(their_sigs), dtype=np.uint8).reshape(n, -1), ctx=b"other").any()


def test_x25519(cb):
    from cryptography.hazmat.primitives import serialization
    from circl_b200 import hybrid
    raw = dict(encoding=serialization.Encoding.Raw, format=serialization.PublicFormat.Raw)
    n = 64
    a, b = _rows(5, n, 32), _rows(6, n, 32)
    pa, pb = hybrid.x25519_keygen(a), hybrid.x25519_keygen(b)
    sab, ok = hybrid.x25519_shared(a, pb)
    assert ok.all()
    for i in range(n):
        ska = ossl_x25519.X25519PrivateKey.from_private_bytes(a[i].tobytes())
        skb = ossl_x25519.X25519PrivateKey.from_private_bytes(b[i].tobytes())
        assert ska.public_key().public_bytes(**raw) == pa[i].tobytes()
        assert skb.public_key().public_bytes(**raw) == pb[i].tobytes()
        assert ska.exchange(skb.public_key()) == sab[i].tobytes()
