"""GPU parity for the samplers and the serialisation leaf methods on their own, through the C ABI (VERDICT r1 row 13 and
"missing" item 9): the vectors the reference embeds in its own tests

  pke/kyber/internal/common/sample_test.go:23-138      DeriveUniform (x=1, y=0), DeriveNoise2/3 (nonce 37)
  sign/mldsa/mldsa65/internal/sample_test.go:12-63     PolyDeriveUniform / LeqEta / LeGamma1 (nonce 30000)

(seed = bytes 0..31 / 0..63 in every vector) and the oracle on random seeds, for every parameter set.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Q = 3329
DQ = 8380417


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def test_kyber_sampler_vectors_of_the_reference(cb, sampler_vectors):
    from circl_b200 import kyber
    seed = np.arange(32, dtype=np.uint8)
    got = kyber.derive_uniform(seed, np.array([[1, 0]], dtype=np.uint8))[0]
    assert got.tolist() == sampler_vectors["kyber_uniform_x1_y0"]
    for eta, key in ((2, "kyber_noise2_nonce37"), (3, "kyber_noise3_nonce37")):
        got = kyber.derive_noise(seed, np.array([37], dtype=np.uint8), eta)[0]
        assert got.tolist() == sampler_vectors[key]


def test_kyber_derive_uniform_vs_oracle_including_four_block_streams(cb):
    import torch
    import oracle
    from circl_b200 import kyber
    rng = np.random.default_rng(168)
    n = 3000  # ~1 % of the streams need a fourth SHAKE128 block
    seeds = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    xy = rng.integers(0, 4, size=(n, 2), dtype=np.uint8)
    got = kyber.derive_uniform(seeds, xy)
    idx = list(range(0, n, 7))
    for i in idx:
        assert np.array_equal(got[i], oracle.kyber_derive_uniform(seeds[i].tobytes(), int(xy[i, 0]), int(xy[i, 1]))), i
    assert int(got.min()) >= 0 and int(got.max()) < Q
    d = kyber.derive_uniform(torch.from_numpy(seeds).cuda(), torch.from_numpy(xy).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), got)
    shared = kyber.derive_uniform(seeds[0], xy[:64].copy())  # one seed for the whole batch: a matrix of one key
    for i in range(64):
        assert np.array_equal(shared[i], oracle.kyber_derive_uniform(seeds[0].tobytes(), int(xy[i, 0]), int(xy[i, 1])))


@pytest.mark.parametrize("eta", [2, 3])
def test_kyber_derive_noise_vs_oracle(cb, eta):
    import oracle
    from circl_b200 import kyber
    rng = np.random.default_rng(eta)
    n = 333
    seeds = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    nonces = rng.integers(0, 256, size=(n,), dtype=np.uint8)
    got = kyber.derive_noise(seeds, nonces, eta)
    for i in range(0, n, 3):
        assert np.array_equal(got[i], oracle.kyber_derive_noise(seeds[i].tobytes(), int(nonces[i]), eta)), i
    assert int(np.abs(got).max()) <= eta


def test_kyber_pack_compress_roundtrips_vs_oracle(cb):
    # poly_test.go: Pack/Unpack and Compress/Decompress against the generic code
    import oracle
    from circl_b200 import kyber
    rng = np.random.default_rng(12)
    n = 67
    p = rng.integers(0, Q, size=(n, 256), dtype=np.int64).astype(np.int16)
    packed = kyber.pack(p)
    for i in range(n):
        assert packed[i].tobytes() == oracle.kyber_pack(p[i])
    assert np.array_equal(kyber.unpack(packed), p)
    raw = rng.integers(0, 256, size=(n, 384), dtype=np.uint8)  # Unpack takes any 12-bit fields (no reduction)
    un = kyber.unpack(raw)
    for i in range(0, n, 5):
        assert np.array_equal(un[i], oracle.kyber_unpack(raw[i].tobytes()))
    for d in (1, 4, 5, 10, 11):
        c = kyber.compress(p, d)
        assert c.shape == (n, 32 * d)
        for i in range(n):
            assert c[i].tobytes() == oracle.kyber_compress(p[i], d), (d, i)
        bits = rng.integers(0, 256, size=(n, 32 * d), dtype=np.uint8)
        dec = kyber.decompress(bits, d)
        for i in range(0, n, 3):
            assert np.array_equal(dec[i], oracle.kyber_decompress(bits[i].tobytes(), d)), (d, i)


def test_dilithium_sampler_vectors_of_the_reference(cb, sampler_vectors):
    from circl_b200 import dilithium
    seed32, seed64 = np.arange(32, dtype=np.uint8), np.arange(64, dtype=np.uint8)
    nonce = np.array([30000], dtype=np.uint16)
    assert dilithium.derive_uniform(seed32, nonce)[0].tolist() == sampler_vectors["dil_uniform_nonce30000"]
    # the reference compares these two after p.Normalize() (mode3/internal/params_test.go:38-39, :99-100); so do we,
    # with the Normalize of the same library
    eta = dilithium.normalize(dilithium.derive_leq_eta(65, seed64, nonce))
    assert eta[0].tolist() == sampler_vectors["dil_leqeta4_nonce30000"]
    g1 = dilithium.normalize(dilithium.derive_le_gamma1(65, seed64, nonce))
    assert g1[0].tolist() == sampler_vectors["dil_legamma1_19_nonce30000"]


@pytest.mark.parametrize("mode", [44, 65, 87])
def test_dilithium_samplers_vs_oracle(cb, mode):
    import torch
    import oracle
    from circl_b200 import dilithium
    rng = np.random.default_rng(mode)
    n = 150
    s32 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    s64 = rng.integers(0, 256, size=(n, 64), dtype=np.uint8)
    nonces = rng.integers(0, 1 << 16, size=(n,), dtype=np.uint16)
    u = dilithium.derive_uniform(s32, nonces)
    e = dilithium.derive_leq_eta(mode, s64, nonces)
    g = dilithium.derive_le_gamma1(mode, s64, nonces)
    ln = {44: 32, 65: 48, 87: 64}[mode]
    ct = rng.integers(0, 256, size=(n, ln), dtype=np.uint8)
    b = dilithium.derive_ball(mode, ct)
    for i in range(0, n, 3):
        assert np.array_equal(u[i], oracle.dil_derive_uniform(s32[i].tobytes(), int(nonces[i]))), i
        assert np.array_equal(e[i], oracle.mldsa_derive_leqeta(mode, s64[i].tobytes(), int(nonces[i]))), i
        assert np.array_equal(g[i], oracle.mldsa_derive_legamma1(mode, s64[i].tobytes(), int(nonces[i]))), i
        assert np.array_equal(b[i], oracle.mldsa_derive_ball(mode, ct[i].tobytes())), i
    tau = {44: 39, 65: 49, 87: 60}[mode]
    assert all(int(np.count_nonzero(row)) == tau for row in b)
    d = dilithium.derive_le_gamma1(mode, torch.from_numpy(s64).cuda(), torch.from_numpy(nonces.view(np.int16)).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint32), g)


def test_dilithium_power2round_and_pack_le16_vs_oracle(cb):
    import oracle
    from circl_b200 import dilithium
    rng = np.random.default_rng(13)
    n = 41
    p = rng.integers(0, DQ, size=(n, 256), dtype=np.int64).astype(np.uint32)
    p[0, :8] = [0, 1, 4095, 4096, 4097, 8191, 8192, DQ - 1]
    a0, a1 = dilithium.power2round(p)
    for i in range(n):
        w0, w1 = oracle.dil_power2round(p[i])
        assert np.array_equal(a0[i], w0) and np.array_equal(a1[i], w1)
    small = rng.integers(0, 16, size=(n, 256), dtype=np.int64).astype(np.uint32)
    packed = dilithium.pack_le16(small)
    for i in range(n):
        assert packed[i].tobytes() == oracle.dil_pack_le16(small[i])
