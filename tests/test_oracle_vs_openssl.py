"""Independent cross-check of the oracle: OpenSSL's ML-KEM / ML-DSA / X25519 (through the `cryptography` package
that ships in this image) against the C restatement of CIRCL -- equal keys from equal seeds, each side decapsulates /
verifies what the other produced.  Neither implementation shares code with the reference or with this repository."""
import hashlib

import pytest

import oracle

mlkem = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.mlkem")
mldsa = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.mldsa")
x25519 = pytest.importorskip("cryptography.hazmat.primitives.asymmetric.x25519")

KEM = {3: ("MLKEM768PrivateKey", "MLKEM768PublicKey"), 4: ("MLKEM1024PrivateKey", "MLKEM1024PublicKey")}
DSA = {44: ("MLDSA44PrivateKey", "MLDSA44PublicKey"), 65: ("MLDSA65PrivateKey", "MLDSA65PublicKey"),
       87: ("MLDSA87PrivateKey", "MLDSA87PublicKey")}


def _h(tag, i, n):
    return hashlib.shake_256(bytes([tag]) + i.to_bytes(4, "little")).digest(n)


@pytest.mark.parametrize("k", list(KEM))
def test_mlkem_against_openssl(k):
    priv_cls, pub_cls = (getattr(mlkem, n) for n in KEM[k])
    for i in range(8):
        seed = _h(1, i, 64)
        ek, dk = oracle.mlkem_keygen(k, seed)
        theirs = priv_cls.from_seed_bytes(seed)
        assert theirs.public_key().public_bytes_raw() == ek
        ct, ss = oracle.mlkem_encaps(k, ek, _h(2, i, 32))
        assert theirs.decapsulate(ct) == ss                                # they decapsulate ours
        a, b = pub_cls.from_public_bytes(ek).encapsulate()
        ss2, ct2 = (a, b) if len(a) == 32 else (b, a)
        assert oracle.mlkem_decaps(k, dk, ct2) == ss2                      # we decapsulate theirs
        bad = bytearray(ct)
        bad[i] ^= 1
        assert theirs.decapsulate(bytes(bad)) == oracle.mlkem_decaps(k, dk, bytes(bad))  # implicit rejection agrees


@pytest.mark.parametrize("mode", list(DSA))
def test_mldsa_against_openssl(mode):
    priv_cls, pub_cls = (getattr(mldsa, n) for n in DSA[mode])
    for i in range(4):
        seed = _h(3, i, 32)
        pk, sk = oracle.mldsa_keygen(mode, seed)
        theirs = priv_cls.from_seed_bytes(seed)
        assert theirs.public_key().public_bytes_raw() == pk
        msg, ctx = _h(4, i, 10 + 30 * i), (b"" if i % 2 == 0 else b"context %d" % i)
        sig, _ = oracle.mldsa_sign(mode, sk, msg, ctx=ctx)
        theirs.public_key().verify(sig, msg, ctx if ctx else None)         # raises InvalidSignature on failure
        assert oracle.mldsa_verify(mode, pk, msg, theirs.sign(msg, ctx if ctx else None), ctx=ctx)
        tampered = bytearray(sig)
        tampered[40] ^= 4
        with pytest.raises(Exception):
            theirs.public_key().verify(bytes(tampered), msg, ctx if ctx else None)
        assert not oracle.mldsa_verify(mode, pk, msg, bytes(tampered), ctx=ctx)


def test_x25519_against_openssl():
    from cryptography.hazmat.primitives import serialization
    raw = dict(encoding=serialization.Encoding.Raw, format=serialization.PublicFormat.Raw)
    for i in range(16):
        a, b = _h(5, i, 32), _h(6, i, 32)
        ska, skb = x25519.X25519PrivateKey.from_private_bytes(a), x25519.X25519PrivateKey.from_private_bytes(b)
        pa, pb = oracle.x25519(a)[0], oracle.x25519(b)[0]
        assert ska.public_key().public_bytes(**raw) == pa and skb.public_key().public_bytes(**raw) == pb
        assert ska.exchange(skb.public_key()) == oracle.x25519(a, pb)[0] == oracle.x25519(b, pa)[0]


@pytest.mark.parametrize("name", ["X25519MLKEM768"])
def test_hybrid_composition_rebuilt_from_openssl_parts(name):
    """kem/hybrid has no vectors in the reference: rebuild X25519MLKEM768 (hybrid.go:197-283 over xkem.go:118-183) from
    OpenSSL's ML-KEM-768 and X25519 plus hashlib's SHAKE256, and compare with the oracle's composition."""
    from cryptography.hazmat.primitives import serialization
    raw = dict(encoding=serialization.Encoding.Raw, format=serialization.PublicFormat.Raw)
    for i in range(6):
        seed, eseed = _h(7, i, 64), _h(8, i, 32)
        pk, sk = oracle.hybrid_keygen(name, seed)
        ex = hashlib.shake_256(seed).digest(96)                  # first.SeedSize (64) || second.SeedSize (32)
        m_priv = mlkem.MLKEM768PrivateKey.from_seed_bytes(ex[:64])
        x_sk = hashlib.shake_256(ex[64:]).digest(32)             # xScheme.DeriveKeyPair
        x_priv = x25519.X25519PrivateKey.from_private_bytes(x_sk)
        assert pk == m_priv.public_key().public_bytes_raw() + x_priv.public_key().public_bytes(**raw)
        assert sk[2400:] == x_sk
        ct, ss, rc = oracle.hybrid_encaps(name, pk, eseed)
        assert rc == 0
        es = hashlib.shake_256(eseed).digest(64)                 # first / second encapsulation seeds
        e_priv = x25519.X25519PrivateKey.from_private_bytes(hashlib.shake_256(es[32:]).digest(32))
        assert ct[1088:] == e_priv.public_key().public_bytes(**raw)
        assert ss[32:] == e_priv.exchange(x_priv.public_key())
        assert ss[:32] == m_priv.decapsulate(ct[:1088])
        assert ct[:1088] == oracle.mlkem_encaps(3, pk[:1184], es[:32])[0]
        assert oracle.hybrid_decaps(name, sk, ct) == (ss, 0)


def test_xwing_composition_rebuilt_from_openssl_parts():
    """X-Wing (kem/xwing/xwing.go:47-66,108-130,209-272) from OpenSSL's parts and hashlib's SHA3-256 / SHAKE256."""
    from cryptography.hazmat.primitives import serialization
    raw = dict(encoding=serialization.Encoding.Raw, format=serialization.PublicFormat.Raw)
    for i in range(6):
        seed, eseed = _h(9, i, 32), _h(10, i, 64)
        pk = oracle.xwing_keygen(seed)
        ex = hashlib.shake_256(seed).digest(96)
        m_priv = mlkem.MLKEM768PrivateKey.from_seed_bytes(ex[:64])
        x_priv = x25519.X25519PrivateKey.from_private_bytes(ex[64:])
        pk_x = x_priv.public_key().public_bytes(**raw)
        assert pk == m_priv.public_key().public_bytes_raw() + pk_x
        ct, ss = oracle.xwing_encaps(pk, eseed)
        e_priv = x25519.X25519PrivateKey.from_private_bytes(eseed[32:])
        ct_x = e_priv.public_key().public_bytes(**raw)
        assert ct[1088:] == ct_x
        ss_x, ss_m = e_priv.exchange(x_priv.public_key()), m_priv.decapsulate(ct[:1088])
        assert ss == hashlib.sha3_256(ss_m + ss_x + ct_x + pk_x + b"\\.//^\\").digest()
        assert oracle.xwing_decaps(seed, ct) == ss
