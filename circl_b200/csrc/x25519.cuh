// X25519 on one thread: GF(2^255 - 19) arithmetic and the Montgomery ladder (RFC 7748 section 5).
//
// Mirrors what dh/x25519/key.go:44-56 computes: KeyGen = u([clamp(k)] G) and Shared = u([clamp(k)] P), both in
// canonical little-endian form; the reference reaches the same values with a Joye ladder over a precomputed table
// (curve.go:7-37) and a Montgomery ladder over 64-bit limbs (curve.go:40-75, math/fp25519).
//
// Representation: ten signed limbs of alternately 26 and 25 bits (value = sum v[i] 2^ceil(25.5 i)), products
// accumulated in 64 bits -- on sm_100a each product-accumulate is one IMAD.WIDE.  With ten independent accumulators
// per multiplication the wide-multiply latency is covered inside a single thread, so the kernel runs one scalar
// multiplication per thread and needs no shared memory.
#pragma once
#include <cstdint>

namespace cb200 {
namespace x25519 {

struct Fe {
  int32_t v[10];
};

__device__ __forceinline__ void fe_set(Fe& h, int32_t x) {
  h.v[0] = x;
#pragma unroll
  for (int i = 1; i < 10; i++) h.v[i] = 0;
}
__device__ __forceinline__ void fe_add(Fe& h, const Fe& f, const Fe& g) {
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = f.v[i] + g.v[i];
}
__device__ __forceinline__ void fe_sub(Fe& h, const Fe& f, const Fe& g) {
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = f.v[i] - g.v[i];
}
__device__ __forceinline__ void fe_cswap(Fe& f, Fe& g, uint32_t bit) {
  const int32_t m = -(int32_t)bit;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const int32_t x = m & (f.v[i] ^ g.v[i]);
    f.v[i] ^= x;
    g.v[i] ^= x;
  }
}

// Rounding carries bring every limb back to |v[2i]| <= 2^25, |v[2i+1]| <= 2^24 (plus the small spill of the final step).
__device__ __forceinline__ void fe_carry(Fe& h, int64_t (&t)[10]) {
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int w = (i & 1) ? 25 : 26;
    const int64_t c = (t[i] + ((int64_t)1 << (w - 1))) >> w;
    t[i + 1] += c;
    t[i] -= c << w;
  }
  {
    const int64_t c = (t[9] + ((int64_t)1 << 24)) >> 25;
    t[0] += 19 * c;
    t[9] -= c << 25;
  }
  {
    const int64_t c = (t[0] + ((int64_t)1 << 25)) >> 26;
    t[1] += c;
    t[0] -= c << 26;
  }
#pragma unroll
  for (int i = 0; i < 10; i++) h.v[i] = (int32_t)t[i];
}

// h = f * g.  Inputs may be one addition or subtraction away from a carried value (|v| < 2^27).
// 2^ceil(25.5 i) * 2^ceil(25.5 j) = 2^ceil(25.5 (i + j)) * (2 if i and j are both odd), and 2^255 = 19.
__device__ __forceinline__ void fe_mul(Fe& h, const Fe& f, const Fe& g) {
  int32_t g19[10], f2[10];
#pragma unroll
  for (int i = 0; i < 10; i++) {
    g19[i] = 19 * g.v[i];
    f2[i] = (i & 1) ? 2 * f.v[i] : f.v[i];
  }
  int64_t t[10];
#pragma unroll
  for (int k = 0; k < 10; k++) {
    int64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      const int j = (k - i + 10) % 10;
      const bool wrap = i > k;
      const int32_t a = (j & 1) ? f2[i] : f.v[i];  // doubled only when both indices are odd
      const int32_t b = wrap ? g19[j] : g.v[j];
      acc += (int64_t)a * b;
    }
    t[k] = acc;
  }
  fe_carry(h, t);
}
// h = f^2: every unordered pair of limbs once (55 wide multiplications instead of 100)
__device__ __forceinline__ void fe_sq(Fe& h, const Fe& f) {
  int32_t f19[10], fx2[10], fx4[10];
#pragma unroll
  for (int i = 0; i < 10; i++) {
    f19[i] = 19 * f.v[i];
    fx2[i] = 2 * f.v[i];
    fx4[i] = 4 * f.v[i];
  }
  int64_t t[10];
#pragma unroll
  for (int k = 0; k < 10; k++) {
    int64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 10; i++) {
      const int j = (k - i + 10) % 10;
      if (i > j) continue;
      const bool wrap = i + j >= 10, both_odd = (i & 1) && (j & 1);
      const int coef = (i == j ? 1 : 2) * (both_odd ? 2 : 1);
      const int32_t a = coef == 1 ? f.v[i] : coef == 2 ? fx2[i] : fx4[i];
      const int32_t b = wrap ? f19[j] : f.v[j];
      acc += (int64_t)a * b;
    }
    t[k] = acc;
  }
  fe_carry(h, t);
}
__device__ __forceinline__ void fe_mul_small(Fe& h, const Fe& f, int32_t s) {
  int64_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = (int64_t)f.v[i] * s;
  fe_carry(h, t);
}

// z^(p - 2) by the usual 2^k - 1 addition chain (254 squarings, 11 multiplications)
__device__ __noinline__ void fe_invert(Fe& out, const Fe& z) {
  Fe z2, z9, z11, z5, z10, z20, z50, z100, t;
  auto pow2k = [](Fe& h, const Fe& f, int k) {
    fe_sq(h, f);
#pragma unroll 1
    for (int i = 1; i < k; i++) fe_sq(h, h);
  };
  pow2k(z2, z, 1);
  pow2k(t, z2, 2);
  fe_mul(z9, t, z);
  fe_mul(z11, z9, z2);
  pow2k(t, z11, 1);
  fe_mul(z5, t, z9);  // 2^5 - 1
  pow2k(t, z5, 5);
  fe_mul(z10, t, z5);  // 2^10 - 1
  pow2k(t, z10, 10);
  fe_mul(z20, t, z10);  // 2^20 - 1
  pow2k(t, z20, 20);
  fe_mul(t, t, z20);  // 2^40 - 1
  pow2k(t, t, 10);
  fe_mul(z50, t, z10);  // 2^50 - 1
  pow2k(t, z50, 50);
  fe_mul(z100, t, z50);  // 2^100 - 1
  pow2k(t, z100, 100);
  fe_mul(t, t, z100);  // 2^200 - 1
  pow2k(t, t, 50);
  fe_mul(t, t, z50);  // 2^250 - 1
  pow2k(t, t, 5);
  fe_mul(out, t, z11);  // 2^255 - 21
}

// 32 little-endian bytes (as 8 words) -> limbs; bit 255 is dropped (key.go:50: validPk[31] &= 127)
__device__ __forceinline__ void fe_frombytes(Fe& h, const uint32_t (&w)[8]) {
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const int off = (51 * i + 1) / 2, width = (i & 1) ? 25 : 26, wi = off >> 5, sh = off & 31;
    uint32_t x = w[wi] >> sh;
    if (sh + width > 32) x |= w[wi + 1] << (32 - sh);
    h.v[i] = (int32_t)(x & ((1u << width) - 1));
  }
}
// canonical encoding (math/fp25519 ToBytes): subtract p exactly when the value is >= p
__device__ __forceinline__ void fe_tobytes(uint32_t (&w)[8], const Fe& f) {
  int32_t h[10];
#pragma unroll
  for (int i = 0; i < 10; i++) h[i] = f.v[i];
  int32_t q = (19 * h[9] + (1 << 24)) >> 25;
#pragma unroll
  for (int i = 0; i < 10; i++) q = (h[i] + q) >> ((i & 1) ? 25 : 26);
  h[0] += 19 * q;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const int wd = (i & 1) ? 25 : 26;
    const int32_t c = h[i] >> wd;
    h[i + 1] += c;
    h[i] -= c << wd;
  }
  h[9] &= (1 << 25) - 1;
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) {
    const int off = (51 * i + 1) / 2, wi = off >> 5, sh = off & 31;
    const uint64_t x = (uint64_t)(uint32_t)h[i] << sh;
    w[wi] |= (uint32_t)x;
    if (wi + 1 < 8) w[wi + 1] |= (uint32_t)(x >> 32);
  }
}

// the five points of small order of curve.go:89-125, canonical form; key.go:25-32 rejects them after reduction mod p
__device__ __forceinline__ bool is_low_order(const uint32_t (&c)[8]) {
  const uint32_t o8a[8] = {0x7c7aebe0u, 0xaeb8413bu, 0xfae35616u, 0x6ac49ff1u, 0xeb8d09dau, 0xfdb1329cu, 0x16056286u, 0x00b8495fu};
  const uint32_t o8b[8] = {0xbc959c5fu, 0x248c50a3u, 0x55b1d0b1u, 0x5bef839cu, 0xc45c4404u, 0x868e1c58u, 0xdd4e22d8u, 0x57119fd0u};
  uint32_t rest = 0, da = 0, db = 0, dm = (c[0] ^ 0xffffffecu) | (c[7] ^ 0x7fffffffu);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if (i > 0) rest |= c[i];
    da |= c[i] ^ o8a[i];
    db |= c[i] ^ o8b[i];
    if (i > 0 && i < 7) dm |= ~c[i];
  }
  const bool zero_or_one = rest == 0 && c[0] <= 1;
  return zero_or_one || da == 0 || db == 0 || dm == 0;
}

// out = u([clamp(k)] P) as 8 words; returns false when P is of small order (the result is then all zero).
__device__ __forceinline__ bool scalarmult(uint32_t (&out)[8], const uint32_t (&kin)[8], const uint32_t (&pin)[8]) {
  uint32_t k[8];
#pragma unroll
  for (int i = 0; i < 8; i++) k[i] = kin[i];
  k[0] &= 0xfffffff8u;
  k[7] = (k[7] & 0x7fffffffu) | 0x40000000u;  // key.go:17-22
  Fe x1, x2, z2, x3, z3;
  fe_frombytes(x1, pin);
  uint32_t canon[8];
  fe_tobytes(canon, x1);
  const bool ok = !is_low_order(canon);
  fe_set(x2, 1);
  fe_set(z2, 0);
  x3 = x1;
  fe_set(z3, 1);
  uint32_t swap = 0;
#pragma unroll 1
  for (int s = 254; s >= 0; s--) {
    uint32_t word = k[0];
#pragma unroll
    for (int q = 1; q < 8; q++) word = ((s >> 5) == q) ? k[q] : word;
    const uint32_t bit = (word >> (s & 31)) & 1;
    swap ^= bit;
    fe_cswap(x2, x3, swap);
    fe_cswap(z2, z3, swap);
    swap = bit;
    Fe a, b, c, d, aa, bb, e, da, cb, t;
    fe_add(a, x2, z2);
    fe_sub(b, x2, z2);
    fe_add(c, x3, z3);
    fe_sub(d, x3, z3);
    fe_sq(aa, a);
    fe_sq(bb, b);
    fe_mul(da, d, a);
    fe_mul(cb, c, b);
    fe_sub(e, aa, bb);
    fe_add(t, da, cb);
    fe_sq(x3, t);
    fe_sub(t, da, cb);
    fe_sq(t, t);
    fe_mul(z3, x1, t);
    fe_mul(x2, aa, bb);
    fe_mul_small(t, e, 121665);
    fe_add(t, t, aa);
    fe_mul(z2, e, t);
  }
  fe_cswap(x2, x3, swap);
  fe_cswap(z2, z3, swap);
  fe_invert(z2, z2);
  fe_mul(x2, x2, z2);
  fe_tobytes(out, x2);
  return ok;
}

}  // namespace x25519
}  // namespace cb200
