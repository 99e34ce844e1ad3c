"""Oracle pins for the wire-side callers: X25519 against the reference's own vectors (dh/x25519/key_test.go) and
X-Wing against the draft's test-vector hash (kem/xwing/xwing_test.go:40-83)."""
import hashlib

import pytest

import oracle
from conftest import load_golden


@pytest.fixture(scope="module")
def xv():
    return load_golden("x25519_vectors.json.gz")


def test_rfc7748_kat(xv):
    for v in xv["rfc7748_kat"]:  # key_test.go:23-45
        out, _ = oracle.x25519(bytes.fromhex(v["scalar"]), bytes.fromhex(v["input"]))
        assert out.hex() == v["output"]


def test_rfc7748_iterated(xv):
    for v in xv["rfc7748_times"]:  # key_test.go:53-86
        u = k = bytes([9] + [0] * 31)
        for _ in range(v["times"]):
            r, _ = oracle.x25519(k, u)
            u, k = k, r
        assert k.hex() == v["key"]


def test_wycheproof(xv):
    for v in xv["wycheproof"]:  # key_test.go:104-141
        out, ok = oracle.x25519(bytes.fromhex(v["private"]), bytes.fromhex(v["public"]))
        assert out.hex() == v["shared"], v["tcId"]
        assert ok or v["result"] == "acceptable", v["tcId"]


def test_keygen_equals_shared_with_base_point():
    for i in range(64):  # key_test.go:88-101
        k = hashlib.shake_256(b"x%d" % i).digest(32)
        assert oracle.x25519(k)[0] == oracle.x25519(k, bytes([9] + [0] * 31))[0]


def xwing_transcript(derive, encaps, decaps):
    """kem/xwing/xwing_test.go:40-76 with pluggable primitives; returns the SHAKE128 of the formatted vectors."""
    stream = hashlib.shake_128(b"").digest(3 * 96)
    w = []

    def write_hex(prefix, val):
        hx = val.hex()
        if len(prefix) + len(hx) + 5 < 74:
            w.append("%s     %s\n" % (prefix, hx))
            return
        w.append(prefix + "\n")
        while hx:
            w.append("  " + hx[:72] + "\n")
            hx = hx[72:]

    for i in range(3):
        seed, eseed = stream[96 * i:96 * i + 32], stream[96 * i + 32:96 * i + 96]
        write_hex("seed", seed)
        pk = derive(seed)
        write_hex("sk", seed)
        write_hex("pk", pk)
        write_hex("eseed", eseed)
        ct, ss = encaps(pk, eseed)
        write_hex("ct", ct)
        write_hex("ss", ss)
        assert decaps(seed, ct) == ss
        w.append("\n")
    return hashlib.shake_128("".join(w).encode()).digest(32).hex()


def test_xwing_draft_vectors(xv):
    assert xwing_transcript(oracle.xwing_keygen, oracle.xwing_encaps, oracle.xwing_decaps) == xv["xwing_vectors_shake128"]


@pytest.mark.parametrize("name", list(oracle.HYBRID_IDS))
def test_hybrid_roundtrip_and_low_order(name):
    pksz, sksz, ctsz = oracle.hybrid_sizes(name)
    pk, sk = oracle.hybrid_keygen(name, hashlib.shake_256(name.encode()).digest(64))
    assert (len(pk), len(sk)) == (pksz, sksz)
    ct, ss, rc = oracle.hybrid_encaps(name, pk, bytes(range(32)))
    assert rc == 0 and len(ct) == ctsz and oracle.hybrid_decaps(name, sk, ct) == (ss, 0)
    # kem/hybrid/xkem_test.go:19-68: a low-order X25519 share is kem.ErrPubKey on both sides
    low = bytes.fromhex("e0eb7a7c3b41b8ae1656e3faf19fc46ada098deb9c32b1fd866205165f49b800")
    x_first = name != "X25519MLKEM768"
    bad_pk = low + pk[32:] if x_first else pk[:-32] + low
    bad_ct = low + ct[32:] if x_first else ct[:-32] + low
    assert oracle.hybrid_encaps(name, bad_pk, bytes(32))[2] == 1
    assert oracle.hybrid_decaps(name, sk, bad_ct)[1] == 1


def test_device_x25519_limb_arithmetic_on_host(tmp_path):
    """csrc/x25519.cuh compiled as plain C++ (CUDA qualifiers defined away) against the oracle: ladder, squaring,
    inversion, canonical encoding and the small-order test, including non-canonical and small-order inputs."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    oracle.build()
    exe = str(tmp_path / "x25519_limbs")
    r = subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "test_x25519_limbs.cpp"), "-o", exe,
                        os.path.join(root, "oracle", "liboracle.so"), "-Wl,-rpath," + os.path.join(root, "oracle")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "bad=0" in r.stdout, r.stdout
