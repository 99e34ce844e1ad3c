"""Pins oracle/keccak.c against the reference's Keccak fixtures (CPU only)."""
import hashlib

import oracle


def test_permutation_of_zero_state(sampler_vectors):
    # simd/keccakf1600/f1600x_test.go:9-19
    assert oracle.keccak_f1600([0] * 25) == sampler_vectors["keccak_f1600_of_zero"]


def test_keccak_kats(keccak_kats):
    # internal/sha3/sha3_test.go:55 (KeccakCodePackage ShortMsgKATs), byte-aligned subset
    fns = {
        "SHA3-256": lambda m, n: oracle.sha3_256(m),
        "SHA3-512": lambda m, n: oracle.sha3_512(m),
        "SHAKE128": oracle.shake128,
        "SHAKE256": oracle.shake256,
    }
    total = 0
    for alg, fn in fns.items():
        for kat in keccak_kats[alg]:
            msg = bytes.fromhex(kat["message"])
            want = bytes.fromhex(kat["digest"])
            assert fn(msg, len(want)) == want, (alg, kat["length"])
            total += 1
    assert total > 100


def test_against_hashlib_multiblock():
    for n in (0, 1, 71, 72, 73, 135, 136, 137, 167, 168, 169, 1184, 1568, 4032, 7000):
        m = bytes((i * 7 + n) & 0xFF for i in range(n))
        assert oracle.sha3_256(m) == hashlib.sha3_256(m).digest()
        assert oracle.sha3_512(m) == hashlib.sha3_512(m).digest()
        assert oracle.shake128(m, 700) == hashlib.shake_128(m).digest(700)
        assert oracle.shake256(m, 700) == hashlib.shake_256(m).digest(700)
