"""CPU check of the arithmetic identity behind compress_any (circl_b200/csrc/mlkem.cu): for every 16-bit representative
x' and every d the reference uses, floor((x' 2^d + 1664) M / 2^40) mod 2^d with M = ceil(2^40 / q) equals the
reference's Compress_q (pke/kyber/internal/common/poly.go:262-328) of x' mod q."""
Q = 3329
M = 330282857


def ref_compress(x, d):
    v = (x << d) + Q // 2
    if d <= 5:
        return ((v * 315) >> 20) & ((1 << d) - 1)       # poly.go:270-290
    return (((v * 20642679) >> 32) >> 4) & ((1 << d) - 1)  # poly.go:291-328 (mul-high by 20642679, then >> 4)


def test_constant():
    assert M == -(-(1 << 40) // Q) and M < (1 << 32)
    assert (M * Q - (1 << 40)) * (1 << 28) < (1 << 40)  # the error term stays below one unit for every v < 2^28


def test_every_representative_and_width():
    for d in (4, 5, 10, 11):
        for xp in range(1 << 16):
            v = (xp << d) + Q // 2
            assert v < (1 << 28)
            assert (((v * M) >> 32) >> 8) & ((1 << d) - 1) == ref_compress(xp % Q, d), (d, xp)
