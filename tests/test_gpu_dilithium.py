"""GPU parity: circl_b200 Dilithium ring kernels vs the oracle (bit-exact), through the C ABI.
Mirrors sign/internal/dilithium/ntt_test.go:11 (accelerated == generic, unnormalised) and poly_test.go:5-125."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Q = 8380417


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def rand_lt(rng, n, bound):
    return rng.integers(0, bound, size=(n, 256), dtype=np.int64).astype(np.uint32)


@pytest.mark.parametrize("n", [1, 3, 16, 17, 1000, (1 << 15) + 5])
def test_ntt_matches_generic(cb, n):
    import oracle
    from circl_b200 import dilithium as dl
    rng = np.random.default_rng(n)
    p = rand_lt(rng, n, 2 * Q)                    # nttGeneric precondition: coefficients < 2q
    assert np.array_equal(dl.ntt_(p.copy()), oracle.dil_ntt(p))
    assert np.array_equal(dl.inv_ntt_(p.copy()), oracle.dil_invntt(p))
    anyv = rand_lt(rng, n, 1 << 32)               # wrap-around semantics on arbitrary uint32 too
    assert np.array_equal(dl.ntt_(anyv.copy()), oracle.dil_ntt(anyv))
    assert np.array_equal(dl.inv_ntt_(anyv.copy()), oracle.dil_invntt(anyv))


def test_device_pointers_roundtrip(cb):
    import torch
    import oracle
    from circl_b200 import dilithium as dl
    rng = np.random.default_rng(5)
    n = 4099
    p = rand_lt(rng, n, Q)
    d = torch.from_numpy(p.view(np.int32)).cuda()
    dl.ntt_(d)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), oracle.dil_ntt(p))
    dl.reduce_le2q(d, out=d)
    dl.inv_ntt_(d)
    dl.normalize(d, out=d)
    want = (p.astype(object) * (1 << 32)) % Q     # ntt_test.go:25: InvNTT(NTT(p)) = R p
    assert np.array_equal(d.cpu().numpy().view(np.uint32).astype(object), want)


def test_mulhat_dot_elementwise_exceeds(cb):
    import oracle
    from circl_b200 import dilithium as dl
    rng = np.random.default_rng(9)
    n = 301
    a, b = rand_lt(rng, n, 2 * Q), rand_lt(rng, n, 2 * Q)
    assert np.array_equal(dl.mul_hat(a, b), oracle.dil_mulhat(a, b))
    av, bv = rand_lt(rng, n * 5, 2 * Q).reshape(n, 5, 256), rand_lt(rng, n * 5, 2 * Q).reshape(n, 5, 256)
    want = np.zeros((n, 256), dtype=np.uint32)
    for j in range(5):
        want = (want + oracle.dil_mulhat(av[:, j], bv[:, j])).astype(np.uint32)
    assert np.array_equal(dl.poly_dot_hat(av, bv, 5), want)
    x = rand_lt(rng, n, 1 << 32)
    assert np.array_equal(dl.add(a, b), oracle.dil_poly_op(0, a, b))
    assert np.array_equal(dl.sub(a, b), oracle.dil_poly_op(1, a, b))
    assert np.array_equal(dl.reduce_le2q(x), oracle.dil_poly_op(2, x))
    assert np.array_equal(dl.normalize(x), oracle.dil_poly_op(3, x))
    assert np.array_equal(dl.normalize_assuming_le2q(a), oracle.dil_poly_op(4, a))
    assert np.array_equal(dl.mul_by_2_to_d(a), (a << 13).astype(np.uint32))
    norm = rand_lt(rng, n, Q)
    norm[::7, 3] = (Q - 1) // 2   # worst-case centred norm
    for bound in (1, 261888 - 196, (1 << 19) - 196, (Q - 1) // 2, (Q - 1) // 2 + 1):
        want_flags = np.array([oracle.dil_exceeds(norm[i], bound) for i in range(n)], dtype=np.uint8)
        assert np.array_equal(dl.exceeds(norm, bound), want_flags)
