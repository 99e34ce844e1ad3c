"""The input bounds of the low-format fast path (circl_b200/csrc/kyber.cuh: kFwdBound, kInvBound) re-derived on the CPU.

The fast path evaluates the reference's int16 arithmetic in 32-bit registers without wrap-around, so it may only run
where no intermediate value of nttGeneric / invNTTGeneric (pke/kyber/internal/common/ntt.go:60-193) can leave int16.
An interval analysis over the exact butterfly and lazy-Barrett schedule (the reference's InvNTTReductions table, pinned in
tests/golden) gives the largest admissible |input|; the constants in the header must not exceed it, and a randomised
search confirms that the analysis is not vacuous (inputs at the bound never overflow in the exact model)."""
import re

import numpy as np

Q = 3329
HDR = __file__.rsplit("/tests/", 1)[0] + "/circl_b200/csrc/kyber.cuh"


def header_bounds():
    m = re.search(r"constexpr int kFwdBound = (\d+), kInvBound = (\d+);", open(HDR).read())
    return int(m.group(1)), int(m.group(2))


def mont_bound(b):
    # |montReduce(zeta * b)| = |(zeta b - m q) / 2^16| with |zeta| <= q - 1 and |m| <= 2^15
    return (b * (Q - 1)) // 65536 + 1665


def forward_ok(B):
    p = [B] * 256
    l = 128
    while l >= 2:
        for off in range(0, 256 - l, 2 * l):
            for j in range(off, off + l):
                if p[j + l] > 32767:
                    return False
                t = mont_bound(p[j + l])
                if p[j] + t > 32767:
                    return False
                p[j] = p[j + l] = p[j] + t
        l >>= 1
    return True


def inverse_ok(B, table):
    p, r, l = [B] * 256, 0, 2
    while l < 256:
        for off in range(0, 256 - l, 2 * l):
            for j in range(off, off + l):
                s = p[j] + p[j + l]          # bounds both p[j] + p[j+l] and p[j+l] - p[j]
                if s > 32767:
                    return False
                p[j], p[j + l] = s, mont_bound(s)
        while True:
            i = table[r]
            r += 1
            if i < 0:
                break
            p[i] = Q                          # barrettReduce returns a value in [0, q]
        l <<= 1
    return True


def test_header_bounds_are_admissible_and_tight(sampler_vectors):
    table = sampler_vectors["kyber_invntt_reductions"]
    fwd, inv = header_bounds()
    assert fwd >= Q and inv >= Q                     # the reference's own contract |c| <= q is inside the fast range
    assert forward_ok(fwd) and inverse_ok(inv, table)
    assert not forward_ok(fwd + 1) and not inverse_ok(inv + 1, table)   # the header uses the largest admissible values


def test_exact_model_never_wraps_inside_the_bounds(sampler_vectors):
    """Exact int arithmetic (no wrap-around) of both transforms on inputs at and inside the bounds stays within int16 at
    every intermediate step -- the property the fast path relies on -- and equals the oracle (which wraps like Go)."""
    import oracle
    table = sampler_vectors["kyber_invntt_reductions"]
    zetas = [int(z) for z in oracle.kyber_zetas()]
    fwd, inv = header_bounds()

    def mont(x):
        m = (x * 62209) & 0xFFFF
        m -= 0x10000 if m & 0x8000 else 0
        return (x - m * Q) >> 16

    def chk(v):
        assert -32768 <= v <= 32767
        return v

    def ntt(p):
        p, k, l = list(p), 0, 128
        while l >= 2:
            for off in range(0, 256 - l, 2 * l):
                k += 1
                for j in range(off, off + l):
                    t = chk(mont(zetas[k] * p[j + l]))
                    p[j + l] = chk(p[j] - t)
                    p[j] = chk(p[j] + t)
            l >>= 1
        return p

    def invntt(p):
        p, k, r, l = list(p), 127, 0, 2
        while l < 256:
            for off in range(0, 256 - l, 2 * l):
                mz = zetas[k]
                k -= 1
                for j in range(off, off + l):
                    t = chk(p[j + l] - p[j])
                    p[j] = chk(p[j] + p[j + l])
                    p[j + l] = chk(mont(mz * t))
            while True:
                i = table[r]
                r += 1
                if i < 0:
                    break
                p[i] = chk(p[i] - ((p[i] * 20159) >> 26) * Q)
            l <<= 1
        return [chk(mont(1441 * x)) for x in p]

    rng = np.random.default_rng(5)
    for bound, model, ref in ((fwd, ntt, oracle.kyber_ntt), (inv, invntt, oracle.kyber_invntt)):
        cases = [np.full(256, bound), np.full(256, -bound), rng.choice([-bound, bound], size=256)]
        cases += [rng.integers(-bound, bound + 1, size=256) for _ in range(20)]
        for c in cases:
            got = model([int(x) for x in c])
            assert got == ref(c.astype(np.int16)[None, :])[0].astype(int).tolist()
