"""GPU parity for the Keccak surface on its own (SURVEY.md 8(a) "Keccak"), through the C ABI:

  cb200_keccak_f1600  vs  simd/keccakf1600/f1600x_test.go:9-19 (permutation of the zero state, every instance of the
                          batch like every lane of StateX4), the oracle on random states, the 12-round turbo variant
  cb200_sha3          vs  internal/sha3/testdata/keccakKats.json.deflate (the short-message KATs held by the reference),
                          hashlib and the oracle at the block-boundary lengths
"""
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


@pytest.mark.parametrize("n", [1, 4, 127, 128, 129, 5000])
def test_permutation_of_zero_state(cb, sampler_vectors, n):
    # f1600x_test.go:9-19: every interleaved instance of a zero StateX4 permutes to the same known state
    from circl_b200 import keccak
    want = np.array(sampler_vectors["keccak_f1600_of_zero"], dtype=np.uint64)
    st = np.zeros((n, 25), dtype=np.uint64)
    keccak.permute_(st)
    assert np.array_equal(st, np.tile(want, (n, 1)))


def test_permutation_random_states_host_and_device(cb):
    import torch
    import oracle
    from circl_b200 import keccak
    rng = np.random.default_rng(1600)
    n = 777
    st = rng.integers(0, 1 << 63, size=(n, 25), dtype=np.uint64) * 2 + rng.integers(0, 2, size=(n, 25), dtype=np.uint64)
    want = np.array([oracle.keccak_f1600([int(x) for x in row]) for row in st], dtype=np.uint64)
    got = keccak.permute_(st.copy())
    assert np.array_equal(got, want)
    d = torch.from_numpy(st.view(np.int64)).cuda()
    keccak.permute_(d)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), want)
    # turbo: rounds 12..23 only (internal/sha3/keccakf.go:12)
    want_t = np.array([oracle.keccak_f1600_turbo([int(x) for x in row]) for row in st[:64]], dtype=np.uint64)
    assert np.array_equal(keccak.permute_(st[:64].copy(), turbo=True), want_t)


def test_keccak_kats_of_the_reference(cb, keccak_kats):
    # internal/sha3/sha3_test.go TestKeccakKats: byte-aligned messages of the four functions on this path
    from circl_b200 import keccak
    fns = {"SHA3-256": lambda m, n: keccak.sha3_256(m), "SHA3-512": lambda m, n: keccak.sha3_512(m),
           "SHAKE128": keccak.shake128, "SHAKE256": keccak.shake256}
    checked = 0
    for name, fn in fns.items():
        groups = {}
        for v in keccak_kats[name]:
            msg, want = bytes.fromhex(v["message"]), bytes.fromhex(v["digest"])
            groups.setdefault((len(msg), len(want)), []).append((msg, want))
        for (ln, outlen), vs in groups.items():
            msgs = np.frombuffer(b"".join(m for m, _ in vs), dtype=np.uint8).reshape(len(vs), ln)
            out = fn(msgs, outlen)
            for (_, want), o in zip(vs, out):
                assert o.tobytes() == want, (name, ln)
                checked += 1
    assert checked > 100


@pytest.mark.parametrize("inlen", [0, 1, 7, 8, 71, 72, 73, 135, 136, 137, 167, 168, 169, 500])
def test_sponges_against_hashlib(cb, inlen):
    import torch
    from circl_b200 import keccak
    rng = np.random.default_rng(inlen)
    n = 97
    msgs = rng.integers(0, 256, size=(n, inlen), dtype=np.uint8)
    for got, ref in ((keccak.sha3_256(msgs), lambda m: hashlib.sha3_256(m).digest()),
                     (keccak.sha3_512(msgs), lambda m: hashlib.sha3_512(m).digest()),
                     (keccak.shake128(msgs, 400), lambda m: hashlib.shake_128(m).digest(400)),
                     (keccak.shake256(msgs, 33), lambda m: hashlib.shake_256(m).digest(33))):
        for i in (0, 1, n - 1):
            assert got[i].tobytes() == ref(msgs[i].tobytes())
    if inlen:
        d = torch.from_numpy(msgs).cuda()
        out = keccak.shake256(d, 200)
        torch.cuda.synchronize()
        assert out[5].cpu().numpy().tobytes() == hashlib.shake_256(msgs[5].tobytes()).digest(200)
