"""GPU parity: circl_b200 Kyber ring kernels vs the oracle (bit-exact), through the C ABI.

Mirrors the reference's differential tests
  pke/kyber/internal/common/ntt_test.go:49,64   (accelerated path == nttGeneric / invNTTGeneric)
  pke/kyber/internal/common/poly_test.go:135-263 (Add/Sub/MulHat/BarrettReduce/Normalize)
with the stricter bar that the inverse NTT is compared *unnormalised* too.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

Q = 3329


@pytest.fixture(scope="module")
def cb():
    import circl_b200
    circl_b200.init(0)
    yield circl_b200
    circl_b200.shutdown()


def rand_abs_le_q(rng, n):
    # RandAbsLeQ, ntt_test.go:41-47
    return rng.integers(-Q, Q, size=(n, 256), dtype=np.int64).astype(np.int16)


def rand_any(rng, n):
    return rng.integers(-32768, 32768, size=(n, 256), dtype=np.int64).astype(np.int16)


@pytest.mark.parametrize("n", [1, 3, 16, 17, 1000, (1 << 16) + 5])
@pytest.mark.parametrize("gen", [rand_abs_le_q, rand_any])
def test_ntt_host_pointers(cb, n, gen):
    import oracle
    from circl_b200 import kyber
    rng = np.random.default_rng(n)
    p = gen(rng, n)
    want = oracle.kyber_ntt(p)
    got = kyber.ntt_(p.copy())
    assert np.array_equal(got, want)
    want_inv = oracle.kyber_invntt(p)
    got_inv = kyber.inv_ntt_(p.copy())
    assert np.array_equal(got_inv, want_inv)


def test_ntt_device_pointers_and_roundtrip(cb):
    import torch
    import oracle
    from circl_b200 import kyber
    rng = np.random.default_rng(99)
    n = 4099
    p = rand_abs_le_q(rng, n)
    d = torch.from_numpy(p).cuda()
    kyber.ntt_(d)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy(), oracle.kyber_ntt(p))
    # ntt_test.go:83 TestNTT: InvNTT(NTT(p)) == R*p
    kyber.barrett_reduce(d, out=d)
    kyber.inv_ntt_(d)
    kyber.normalize(d, out=d)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().astype(np.int64), (p.astype(np.int64) * 65536) % Q)


@pytest.mark.parametrize("n", [1, 5, 4096 + 3])
def test_mulhat_and_dot(cb, n):
    import oracle
    from circl_b200 import kyber
    rng = np.random.default_rng(7 + n)
    a, b = rand_abs_le_q(rng, n), rand_abs_le_q(rng, n)
    assert np.array_equal(kyber.mul_hat(a, b), oracle.kyber_mulhat(a, b))
    a2, b2 = rand_any(rng, n), rand_any(rng, n)
    assert np.array_equal(kyber.mul_hat(a2, b2), oracle.kyber_mulhat(a2, b2))
    for k in (2, 3, 4):
        av = rng.integers(-Q, Q, size=(n, k, 256)).astype(np.int16)
        bv = rng.integers(-Q, Q, size=(n, k, 256)).astype(np.int16)
        assert np.array_equal(kyber.poly_dot_hat(av, bv, k), oracle.kyber_dot(av, bv, k))


def test_elementwise(cb):
    import oracle
    from circl_b200 import kyber
    rng = np.random.default_rng(5)
    a, b = rand_any(rng, 77), rand_any(rng, 77)
    assert np.array_equal(kyber.add(a, b), (a.astype(np.int32) + b).astype(np.int16))
    assert np.array_equal(kyber.sub(a, b), (a.astype(np.int32) - b).astype(np.int16))
    assert np.array_equal(kyber.barrett_reduce(a), oracle.kyber_barrett(a))
    # Normalize requires x >= -29439 (field.go:66); all int16 values are exercised by barrett first
    assert np.array_equal(kyber.normalize(a), oracle.kyber_normalize(a))
    assert np.array_equal(kyber.to_mont(a), oracle.kyber_tomont(a))


def test_every_int16_value_through_field_ops(cb):
    import oracle
    from circl_b200 import kyber
    allv = np.arange(-32768, 32768, dtype=np.int32).astype(np.int16).reshape(256, 256)
    assert np.array_equal(kyber.barrett_reduce(allv), oracle.kyber_barrett(allv))
    assert np.array_equal(kyber.normalize(allv), oracle.kyber_normalize(allv))
    assert np.array_equal(kyber.to_mont(allv), oracle.kyber_tomont(allv))


def test_empty_batch(cb):
    from circl_b200 import kyber
    e = np.empty((0, 256), dtype=np.int16)
    assert kyber.ntt_(e).shape == (0, 256)


def test_full_size_property_2_20(cb):
    """BASELINE config 2 size: 2^20 polynomials.  Oracle-checked on a strided
    sample, plus the size-independent property InvNTT(NTT(p)) == R*p on all."""
    import torch
    import oracle
    from circl_b200 import kyber
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(1234)
    d = (torch.randint(0, 2 * Q, (n, 256), generator=g, device="cuda", dtype=torch.int32) - Q).to(torch.int16)
    orig = d.clone()
    kyber.ntt_(d)
    idx = torch.arange(0, n, 4099, device="cuda")
    sample_in = orig[idx].cpu().numpy()
    assert np.array_equal(d[idx].cpu().numpy(), oracle.kyber_ntt(sample_in))
    kyber.barrett_reduce(d, out=d)
    kyber.inv_ntt_(d)
    kyber.normalize(d, out=d)
    want = (orig.to(torch.int64) * 65536) % Q
    assert torch.equal(d.to(torch.int64), want)


@pytest.mark.parametrize("inverse,bound", [(False, 13561), (True, 3679)])
def test_ntt_fast_and_general_path_agree_with_oracle(cb, inverse, bound):
    """The NTT kernels choose per warp between the low-format fast path (all coefficients of the warp's four
    polynomials inside [-bound, bound]: no int16 wrap-around can occur in ntt.go:60-193) and the general path.
    Batches that sit at the edge of the range, one past it, and that interleave both kinds -- so that neighbouring
    warps, and polynomials inside one warp, fall on different sides -- must all equal the oracle."""
    import oracle
    from circl_b200 import kyber
    N = 256
    rng = np.random.default_rng(99 + int(inverse))
    n = 4 * 16 * 37 + 3  # ragged: the last warp has idle octets
    cases = {
        "at the bound": rng.integers(-bound, bound + 1, size=(n, N)),
        "extremes of the range": rng.choice([-bound, bound], size=(n, N)),
        "one past the bound": np.where(rng.random((n, N)) < 0.01, rng.choice([-bound - 1, bound + 1], size=(n, N)),
                                       rng.integers(-bound, bound + 1, size=(n, N))),
        "contract": rng.integers(-Q, Q + 1, size=(n, N)),
    }
    mixed = rng.integers(-bound, bound + 1, size=(n, N))
    mixed[::5] = rng.integers(-32768, 32768, size=mixed[::5].shape)       # one polynomial in five anywhere in int16
    mixed[3::64, 255] = bound + 1                                          # a single coefficient decides
    mixed[7::64, 0] = -bound - 1
    mixed[11::64] = -32768
    cases["mixed"] = mixed
    for name, x in cases.items():
        p = x.astype(np.int16)
        want = oracle.kyber_invntt(p) if inverse else oracle.kyber_ntt(p)
        got = (kyber.inv_ntt_ if inverse else kyber.ntt_)(p.copy())
        assert np.array_equal(got, want), name
