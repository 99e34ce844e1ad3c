"""CPU check of the arithmetic behind the IDP.2A matrix-vector products of encrypt_dp_kernel / keygen_dp_kernel /
decrypt_dp_kernel (circl_b200/csrc/mlkem.cu): per degree-2 block and column, with a = (a0, a1) a word of A-hat (plain
residues < 4096) and b = (b0, b1) the operand (|b| < q),

    W0 = [b0.lo, (z b1).lo, b0.hi, (z b1).hi]      W1 = [b1.lo, b0.lo, b1.hi, b0.hi]     (low bytes unsigned, high bytes signed)
    p0 = dp2a.lo(a, W0) + 256 dp2a.hi(a, W0) = a0 b0 + a1 (z b1)
    p1 = dp2a.lo(a, W1) + 256 dp2a.hi(a, W1) = a0 b1 + a1 b0

summed over the K columns in 32 bits, and one montReduce per output coefficient.  Against the reference's own
formulation (mulHatGeneric + PolyDotHat, poly.go:63-100, vec.go:30-37) through the oracle: the results agree modulo q,
which is all that leaves those kernels (Compress / Normalize / message bits follow)."""
import numpy as np

Q = 3329


def s16(x):
    x = np.asarray(x, dtype=np.int64) & 0xFFFF
    return np.where(x & 0x8000, x - 0x10000, x)


def mont(x):  # montReduce, field.go:4-32
    x = np.asarray(x, dtype=np.int64)
    m = s16(x * 62209)
    return (x - m * Q) >> 16


def split(v):  # low byte unsigned, high byte signed: v = lo + 256 hi for any int16 v
    v = np.asarray(v, dtype=np.int64)
    lo = v & 0xFF
    hi = (v - lo) >> 8
    assert np.all((hi >= -128) & (hi <= 127))
    return lo, hi


def test_block_products_equal_mulhat_mod_q():
    import oracle
    rng = np.random.default_rng(11)
    zetas = oracle.kyber_zetas().astype(np.int64)           # Montgomery form, ntt.go:5-15
    for K in (2, 3, 4):
        a = rng.integers(0, 4096, size=(K, 256))             # a word of A-hat / t-hat: any 12-bit value
        b = rng.integers(-Q + 1, Q, size=(K, 256))           # operands after the Montgomery scaling: |b| < q
        want = oracle.kyber_dot(a.astype(np.int16)[None, ...] % Q, b.astype(np.int16)[None, ...], K)[0].astype(np.int64)
        acc = np.zeros(256, dtype=np.int64)
        for j in range(K):
            for blk in range(128):                           # block = coefficients 2 blk, 2 blk + 1
                z = zetas[64 + blk // 2] * (1 if blk % 2 == 0 else -1)   # +zeta for the first block of a quad, -zeta for the second
                a0, a1 = a[j, 2 * blk], a[j, 2 * blk + 1]
                b0, b1 = b[j, 2 * blk], b[j, 2 * blk + 1]
                zb1 = int(mont(z * b1))                      # (z R) b1 R^-1 = z b1, |.| < q
                (b0l, b0h), (b1l, b1h), (zl, zh) = split(b0), split(b1), split(zb1)
                p0 = (a0 * b0l + a1 * zl) + 256 * (a0 * b0h + a1 * zh)
                p1 = (a0 * b1l + a1 * b0l) + 256 * (a0 * b1h + a1 * b0h)
                assert p0 == a0 * b0 + a1 * zb1 and p1 == a0 * b1 + a1 * b0
                acc[2 * blk] += p0
                acc[2 * blk + 1] += p1
        assert np.abs(acc).max() < 2 ** 31                   # the 32-bit accumulators of the kernels never overflow
        got = mont(acc)
        assert np.abs(got).max() <= Q                        # the input bound of the inverse transform's schedule
        assert np.array_equal(got % Q, want % Q), K


def test_operand_constants():
    """The Shoup pairs the kernels scale their operands with: 512 R (= 1441, also the inverse transform's constant) and
    R^2 (= 1353, ToMont) as (zp, kk) with zp = c 2^-16 mod q centred and kk = (zp 2^16 - c) / q."""
    R = (1 << 16) % Q
    assert (512 * R) % Q == 1441 and (R * R) % Q == 1353
    for c, zp in ((1441, 512), (1353, R - Q)):
        assert (zp * 65536 - c) % Q == 0 and (zp * 65536) % Q == c % Q
        kk = (zp * 65536 - c) // Q
        b = np.arange(-32768, 32768, dtype=np.int64)
        n = (kk * b + 32767) >> 16
        assert np.array_equal(zp * b - n * Q, mont(c * b))   # the same integer as montReduce(c b) for every int16 b
