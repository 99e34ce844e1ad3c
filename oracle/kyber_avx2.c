/*
 * oracle/kyber_avx2.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * The second CPU arm of bench.py: ML-KEM Encapsulate the way the reference runs it on amd64, where its hot leaves are
 * not the generic Go code that kyber.c restates but AVX2 assembly:
 *
 *   simd/keccakf1600/f1600x4_amd64.s:9  f1600x4AVX2   four Keccak-f[1600] states, lane i of instance j at a[4 i + j]
 *                                                    (f1600x.go:30-44); used by Mat.Derive through
 *   pke/kyber/internal/common/sample.go:101-187      PolyDeriveUniformX4 (four SHAKE128 streams of matrix A at once)
 *   pke/kyber/internal/common/amd64.s:153,737,1443   nttAVX2, invNttAVX2, mulHatAVX2 (16 int16 lanes per instruction)
 *
 * Everything else of Encapsulate (SHA3-256 / SHA3-512 / the SHAKE256 PRF, CBD, packing, compression) is plain Go in
 * the reference too and is taken from kyber.c / keccak.c unchanged.  This file states the same three leaves with
 * compiler intrinsics.  It is NOT a copy of the assembly (which works on a "tangled" coefficient order, amd64.go:265);
 * it keeps the standard order and moves lanes with unpack/permute around the short-distance layers instead.  Its
 * outputs are checked against kyber.c: Keccak state for state, NTT / MulHat coefficient for coefficient, InvNTT after
 * Normalize (this inverse Barrett-reduces all lanes after layers 3 and 6 instead of the reference's 68 selected
 * coefficients), ciphertexts and shared secrets byte for byte (tests/test_oracle_avx2.py).
 */
#include <immintrin.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

#define N 256
#define Q 3329
#define QINV 62209u /* q^-1 mod 2^16, field.go:12 */

/* ------------------------------------------------------------------ Keccak-f[1600] x 4 */
static uint64_t RC[24];
static int k_ready;
static void k_tables(void) { /* round constants from the LFSR of FIPS 202, as in keccak.c */
  uint8_t lfsr = 1;
  for (int r = 0; r < 24; r++) {
    uint64_t c = 0;
    for (int j = 0; j < 7; j++) {
      if (lfsr & 1) c ^= 1ULL << ((1u << j) - 1);
      lfsr = (uint8_t)((lfsr << 1) ^ ((lfsr & 0x80) ? 0x71 : 0));
    }
    RC[r] = c;
  }
  k_ready = 1;
}
/* rho offsets of lane x + 5 y ((t + 1)(t + 2) / 2 along the walk (x, y) -> (y, 2x + 3y)); literal so that the fully
 * unrolled round below gets immediate shift counts */
static const int RHO[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#define ROL(v, r) ((r) == 0 ? (v) : _mm256_or_si256(_mm256_slli_epi64((v), (r)), _mm256_srli_epi64((v), 64 - (r))))
/* One round from `in` to `out` (two buffers, so rho-pi needs no temporary state): theta's D is applied while the lanes
 * are fetched, and chi is done row by row right after the five lanes of a row have been rotated into place, which keeps
 * about fifteen vectors live -- the sixteen ymm registers, not the stack, carry the round. */
static inline void round_x4(const __m256i *in, __m256i *out, uint64_t rc) {
  __m256i c[5], d[5];
#pragma GCC unroll 5
  for (int x = 0; x < 5; x++)
    c[x] = _mm256_xor_si256(_mm256_xor_si256(_mm256_xor_si256(in[x], in[x + 5]), _mm256_xor_si256(in[x + 10], in[x + 15])), in[x + 20]);
#pragma GCC unroll 5
  for (int x = 0; x < 5; x++) d[x] = _mm256_xor_si256(c[(x + 4) % 5], ROL(c[(x + 1) % 5], 1));
#pragma GCC unroll 5
  for (int yo = 0; yo < 5; yo++) { /* output row yo: lane (xo, yo) comes from lane (x, y) = ((3 yo + xo) mod 5, xo) */
    __m256i t[5];
#pragma GCC unroll 5
    for (int xo = 0; xo < 5; xo++) {
      const int x = (3 * yo + xo) % 5, y = xo;
      t[xo] = ROL(_mm256_xor_si256(in[x + 5 * y], d[x]), RHO[x + 5 * y]);
    }
#pragma GCC unroll 5
    for (int xo = 0; xo < 5; xo++) out[xo + 5 * yo] = _mm256_xor_si256(t[xo], _mm256_andnot_si256(t[(xo + 1) % 5], t[(xo + 2) % 5]));
  }
  out[0] = _mm256_xor_si256(out[0], _mm256_set1_epi64x((long long)rc));
}
/* a: 100 words, lane i of instance j at a[4 i + j] (the StateX4 layout) */
void orc_keccak_f1600_x4(uint64_t *a) {
  if (!k_ready) k_tables();
  __m256i s[25], u[25];
  for (int i = 0; i < 25; i++) s[i] = _mm256_loadu_si256((const __m256i *)(a + 4 * i));
  for (int r = 0; r < 24; r += 2) {
    round_x4(s, u, RC[r]);
    round_x4(u, s, RC[r + 1]);
  }
  for (int i = 0; i < 25; i++) _mm256_storeu_si256((__m256i *)(a + 4 * i), s[i]);
}

/* PolyDeriveUniformX4 (sample.go:101-187): up to four polynomials ps[0..3] (NULL = unused) from SHAKE128(rho || x || y) */
static void derive_uniform_x4(int16_t *ps[4], const uint8_t rho[32], const uint8_t xs[4], const uint8_t ys[4]) {
  uint64_t st[100];
  memset(st, 0, sizeof st);
  for (int w = 0; w < 4; w++) {
    uint64_t v;
    memcpy(&v, rho + 8 * w, 8);
    for (int j = 0; j < 4; j++) st[4 * w + j] = v;
  }
  for (int j = 0; j < 4; j++) {
    st[4 * 4 + j] = (uint64_t)xs[j] | ((uint64_t)ys[j] << 8) | (0x1fULL << 16); /* sample.go:116-119 */
    st[4 * 20 + j] = 0x8000000000000000ULL;                                      /* rate 168 */
  }
  int cnt[4] = {0, 0, 0, 0}, left = 0;
  for (int j = 0; j < 4; j++)
    if (ps[j]) left++;
  while (left) {
    orc_keccak_f1600_x4(st);
    for (int j = 0; j < 4; j++) {
      if (!ps[j] || cnt[j] == N) continue;
      uint8_t buf[168];
      for (int w = 0; w < 21; w++) memcpy(buf + 8 * w, &st[4 * w + j], 8);
      int16_t *p = ps[j];
      int i = cnt[j];
      for (int o = 0; o < 168 && i < N; o += 3) { /* sample.go:157-183 */
        const uint16_t d1 = (uint16_t)(buf[o] | ((buf[o + 1] & 0xf) << 8));
        const uint16_t d2 = (uint16_t)((buf[o + 1] >> 4) | (buf[o + 2] << 4));
        if (d1 < Q) p[i++] = (int16_t)d1;
        if (d2 < Q && i < N) p[i++] = (int16_t)d2;
      }
      cnt[j] = i;
      if (i == N) left--;
    }
  }
}

/* ------------------------------------------------------------------ 16-lane field arithmetic */
static int16_t Z[128], ZQ[128]; /* Zetas and Zetas * q^-1 mod 2^16 */
static int16_t MHZ[8][16], MHZQ[8][16], MHS[8][16]; /* MulHat: zeta, zeta q^-1 and the sign of each block */
static int z_ready;
static void z_tables(void) {
  const int16_t *zt = orc_kyber_zetas();
  for (int i = 0; i < 128; i++) {
    Z[i] = zt[i];
    ZQ[i] = (int16_t)((uint32_t)(uint16_t)zt[i] * QINV);
  }
  /* MulHat: the even-coefficient vector of a group of 32 coefficients holds blocks
   * [0..3 | 8..11 || 4..7 | 12..15] of the group (see mulhat below); block m uses +-Zetas[64 + m/2] (poly.go:70-98) */
  static const int order[16] = {0, 1, 2, 3, 8, 9, 10, 11, 4, 5, 6, 7, 12, 13, 14, 15};
  for (int g = 0; g < 8; g++)
    for (int e = 0; e < 16; e++) {
      const int m = 16 * g + order[e];
      const int16_t z = zt[64 + m / 2];
      MHZ[g][e] = z;
      MHZQ[g][e] = (int16_t)((uint32_t)(uint16_t)z * QINV);
      MHS[g][e] = (m & 1) ? -1 : 1; /* p2 = -montReduce(p2 * zeta), poly.go:87 */
    }
  z_ready = 1;
}
/* montReduce(z * b) per lane (field.go:4-32): hi16(z b) - hi16(int16(z b q^-1) q) */
static inline __m256i mont_zb(__m256i b, __m256i z, __m256i zq) {
  const __m256i m = _mm256_mullo_epi16(b, zq);
  return _mm256_sub_epi16(_mm256_mulhi_epi16(b, z), _mm256_mulhi_epi16(m, _mm256_set1_epi16(Q)));
}
static inline __m256i mont_ab(__m256i a, __m256i b) {
  const __m256i m = _mm256_mullo_epi16(_mm256_mullo_epi16(a, b), _mm256_set1_epi16((int16_t)QINV));
  return _mm256_sub_epi16(_mm256_mulhi_epi16(a, b), _mm256_mulhi_epi16(m, _mm256_set1_epi16(Q)));
}
static inline __m256i barrett16(__m256i x) { /* field.go:45-64: x - ((x * 20159) >> 26) q */
  const __m256i t = _mm256_srai_epi16(_mm256_mulhi_epi16(x, _mm256_set1_epi16(20159)), 10);
  return _mm256_sub_epi16(x, _mm256_mullo_epi16(t, _mm256_set1_epi16(Q)));
}
#define LD(p) _mm256_loadu_si256((const __m256i *)(p))
#define ST(p, v) _mm256_storeu_si256((__m256i *)(p), (v))
static inline __m256i two128(const int16_t *t, int k0, int k1) { /* [t[k0] x 8 | t[k1] x 8] */
  return _mm256_set_m128i(_mm_set1_epi16(t[k1]), _mm_set1_epi16(t[k0]));
}

void orc_kyber_ntt_avx2(int16_t p[N]) { /* ntt.go:60-135, same outputs as orc_kyber_ntt */
  if (!z_ready) z_tables();
  int k = 0;
  for (int l = 128; l >= 16; l >>= 1)
    for (int off = 0; off < N - l; off += 2 * l) {
      ++k;
      const __m256i z = _mm256_set1_epi16(Z[k]), zq = _mm256_set1_epi16(ZQ[k]);
      for (int j = off; j < off + l; j += 16) {
        const __m256i a = LD(p + j), t = mont_zb(LD(p + j + l), z, zq);
        ST(p + j + l, _mm256_sub_epi16(a, t));
        ST(p + j, _mm256_add_epi16(a, t));
      }
    }
  /* l = 8: two blocks of 16 per step; halves of the 128-bit lanes are regrouped into a | b */
  for (int o = 0, kk = 16; o < N; o += 32, kk += 2) {
    const __m256i v0 = LD(p + o), v1 = LD(p + o + 16);
    const __m256i a = _mm256_permute2x128_si256(v0, v1, 0x20), b = _mm256_permute2x128_si256(v0, v1, 0x31);
    const __m256i t = mont_zb(b, two128(Z, kk, kk + 1), two128(ZQ, kk, kk + 1));
    const __m256i a2 = _mm256_add_epi16(a, t), b2 = _mm256_sub_epi16(a, t);
    ST(p + o, _mm256_permute2x128_si256(a2, b2, 0x20));
    ST(p + o + 16, _mm256_permute2x128_si256(a2, b2, 0x31));
  }
  /* l = 4: four blocks of 8 per step; 64-bit groups [a0 a2 | a1 a3] */
  for (int o = 0, kk = 32; o < N; o += 32, kk += 4) {
    const __m256i v0 = LD(p + o), v1 = LD(p + o + 16);
    const __m256i a = _mm256_unpacklo_epi64(v0, v1), b = _mm256_unpackhi_epi64(v0, v1);
    const int16_t zz[16] = {Z[kk], Z[kk], Z[kk], Z[kk], Z[kk + 2], Z[kk + 2], Z[kk + 2], Z[kk + 2],
                            Z[kk + 1], Z[kk + 1], Z[kk + 1], Z[kk + 1], Z[kk + 3], Z[kk + 3], Z[kk + 3], Z[kk + 3]};
    const __m256i z = LD(zz), zq = _mm256_mullo_epi16(z, _mm256_set1_epi16((int16_t)QINV));
    const __m256i t = mont_zb(b, z, zq);
    const __m256i a2 = _mm256_add_epi16(a, t), b2 = _mm256_sub_epi16(a, t);
    ST(p + o, _mm256_unpacklo_epi64(a2, b2));
    ST(p + o + 16, _mm256_unpackhi_epi64(a2, b2));
  }
  /* l = 2: eight blocks of 4 per step; 32-bit groups [a0 a1 a4 a5 | a2 a3 a6 a7] */
  for (int o = 0, kk = 64; o < N; o += 32, kk += 8) {
    const __m256 v0 = _mm256_castsi256_ps(LD(p + o)), v1 = _mm256_castsi256_ps(LD(p + o + 16));
    const __m256i a = _mm256_castps_si256(_mm256_shuffle_ps(v0, v1, 0x88)), b = _mm256_castps_si256(_mm256_shuffle_ps(v0, v1, 0xdd));
    static const int ord[8] = {0, 1, 4, 5, 2, 3, 6, 7};
    int16_t zz[16];
    for (int e = 0; e < 8; e++) zz[2 * e] = zz[2 * e + 1] = Z[kk + ord[e]];
    const __m256i z = LD(zz), zq = _mm256_mullo_epi16(z, _mm256_set1_epi16((int16_t)QINV));
    const __m256i t = mont_zb(b, z, zq);
    const __m256i a2 = _mm256_add_epi16(a, t), b2 = _mm256_sub_epi16(a, t);
    ST(p + o, _mm256_unpacklo_epi32(a2, b2));
    ST(p + o + 16, _mm256_unpackhi_epi32(a2, b2));
  }
}

/* ntt.go:145-193 up to the choice of representatives: all lanes are Barrett-reduced after layers 3 and 6 (bounds 8 q
 * < 2^15 in between), so outputs equal orc_kyber_invntt's modulo q, not bit for bit */
void orc_kyber_invntt_avx2(int16_t p[N]) {
  if (!z_ready) z_tables();
  for (int o = 0, kk = 127; o < N; o += 32, kk -= 8) { /* l = 2 */
    const __m256 v0 = _mm256_castsi256_ps(LD(p + o)), v1 = _mm256_castsi256_ps(LD(p + o + 16));
    const __m256i a = _mm256_castps_si256(_mm256_shuffle_ps(v0, v1, 0x88)), b = _mm256_castps_si256(_mm256_shuffle_ps(v0, v1, 0xdd));
    static const int ord[8] = {0, 1, 4, 5, 2, 3, 6, 7};
    int16_t zz[16];
    for (int e = 0; e < 8; e++) zz[2 * e] = zz[2 * e + 1] = Z[kk - ord[e]];
    const __m256i z = LD(zz), zq = _mm256_mullo_epi16(z, _mm256_set1_epi16((int16_t)QINV));
    const __m256i a2 = _mm256_add_epi16(a, b), b2 = mont_zb(_mm256_sub_epi16(b, a), z, zq);
    ST(p + o, _mm256_unpacklo_epi32(a2, b2));
    ST(p + o + 16, _mm256_unpackhi_epi32(a2, b2));
  }
  for (int o = 0, kk = 63; o < N; o += 32, kk -= 4) { /* l = 4 */
    const __m256i v0 = LD(p + o), v1 = LD(p + o + 16);
    const __m256i a = _mm256_unpacklo_epi64(v0, v1), b = _mm256_unpackhi_epi64(v0, v1);
    const int16_t zz[16] = {Z[kk], Z[kk], Z[kk], Z[kk], Z[kk - 2], Z[kk - 2], Z[kk - 2], Z[kk - 2],
                            Z[kk - 1], Z[kk - 1], Z[kk - 1], Z[kk - 1], Z[kk - 3], Z[kk - 3], Z[kk - 3], Z[kk - 3]};
    const __m256i z = LD(zz), zq = _mm256_mullo_epi16(z, _mm256_set1_epi16((int16_t)QINV));
    const __m256i a2 = _mm256_add_epi16(a, b), b2 = mont_zb(_mm256_sub_epi16(b, a), z, zq);
    ST(p + o, _mm256_unpacklo_epi64(a2, b2));
    ST(p + o + 16, _mm256_unpackhi_epi64(a2, b2));
  }
  for (int o = 0, kk = 31; o < N; o += 32, kk -= 2) { /* l = 8, then the first full reduction */
    const __m256i v0 = LD(p + o), v1 = LD(p + o + 16);
    const __m256i a = _mm256_permute2x128_si256(v0, v1, 0x20), b = _mm256_permute2x128_si256(v0, v1, 0x31);
    const __m256i a2 = barrett16(_mm256_add_epi16(a, b));
    const __m256i b2 = mont_zb(_mm256_sub_epi16(b, a), two128(Z, kk, kk - 1), two128(ZQ, kk, kk - 1));
    ST(p + o, _mm256_permute2x128_si256(a2, b2, 0x20));
    ST(p + o + 16, _mm256_permute2x128_si256(a2, b2, 0x31));
  }
  int k = 15;
  for (int l = 16; l < N; l <<= 1)
    for (int off = 0; off < N - l; off += 2 * l) {
      const __m256i z = _mm256_set1_epi16(Z[k]), zq = _mm256_set1_epi16(ZQ[k]);
      k--;
      for (int j = off; j < off + l; j += 16) {
        const __m256i a = LD(p + j), b = LD(p + j + l);
        __m256i s = _mm256_add_epi16(a, b);
        if (l == 64) s = barrett16(s); /* after layer 6 */
        ST(p + j, s);
        ST(p + j + l, mont_zb(_mm256_sub_epi16(b, a), z, zq));
      }
    }
  const __m256i f = _mm256_set1_epi16(1441), fq = _mm256_set1_epi16((int16_t)(1441u * QINV));
  for (int j = 0; j < N; j += 16) ST(p + j, mont_zb(LD(p + j), f, fq));
}

/* acc += MulHat(a, b) (poly.go:63-100): 32 coefficients per step, even and odd coefficients split by byte shuffles */
static void mulhat_acc_avx2(int16_t acc[N], const int16_t a[N], const int16_t b[N]) {
  if (!z_ready) z_tables();
  const __m256i sh = _mm256_setr_epi8(0, 1, 4, 5, 8, 9, 12, 13, 2, 3, 6, 7, 10, 11, 14, 15, 0, 1, 4, 5, 8, 9, 12, 13, 2, 3, 6, 7,
                                      10, 11, 14, 15);
  for (int g = 0; g < 8; g++) {
    const int o = 32 * g;
    const __m256i a0 = _mm256_shuffle_epi8(LD(a + o), sh), a1 = _mm256_shuffle_epi8(LD(a + o + 16), sh);
    const __m256i b0 = _mm256_shuffle_epi8(LD(b + o), sh), b1 = _mm256_shuffle_epi8(LD(b + o + 16), sh);
    const __m256i ae = _mm256_unpacklo_epi64(a0, a1), ao = _mm256_unpackhi_epi64(a0, a1);
    const __m256i be = _mm256_unpacklo_epi64(b0, b1), bo = _mm256_unpackhi_epi64(b0, b1);
    const __m256i pe = _mm256_add_epi16(_mm256_sign_epi16(mont_zb(mont_ab(ao, bo), LD(MHZ[g]), LD(MHZQ[g])), LD(MHS[g])), mont_ab(ae, be));
    const __m256i po = _mm256_add_epi16(mont_ab(ae, bo), mont_ab(ao, be));
    ST(acc + o, _mm256_add_epi16(LD(acc + o), _mm256_unpacklo_epi16(pe, po)));
    ST(acc + o + 16, _mm256_add_epi16(LD(acc + o + 16), _mm256_unpackhi_epi16(pe, po)));
  }
}
void orc_kyber_mulhat_avx2(int16_t p[N], const int16_t a[N], const int16_t b[N]) {
  int16_t acc[N];
  memset(acc, 0, sizeof acc);
  mulhat_acc_avx2(acc, a, b);
  memcpy(p, acc, sizeof acc);
}

/* ------------------------------------------------------------------ ML-KEM Encapsulate on these leaves */
static int du_of(int k) { return k == 4 ? 11 : 10; }
static int dv_of(int k) { return k == 4 ? 5 : 4; }
static int eta1_of(int k) { return k == 2 ? 3 : 2; }

int orc_mlkem_encaps_avx2(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *ek, const uint8_t m[32]) {
  /* UnmarshalBinaryPublicKey (kyber.go:247-263, cpapke.go:45-63) */
  int16_t th[4 * N], aT[16 * N];
  uint8_t chk[384];
  int bad = 0;
  for (int i = 0; i < k; i++) {
    orc_kyber_unpack(th + i * N, ek + 384 * i);
    orc_kyber_normalize(th + i * N);
    orc_kyber_pack(chk, th + i * N);
    if (memcmp(chk, ek + 384 * i, 384)) bad = 1;
  }
  if (bad) return -1;
  const uint8_t *rho = ek + 384 * k;
  for (int e = 0; e < k * k; e += 4) { /* Mat.Derive, transposed, four entries at a time (mat.go:31-74) */
    int16_t *ps[4];
    uint8_t xs[4], ys[4];
    for (int j = 0; j < 4; j++) {
      const int idx = e + j;
      ps[j] = idx < k * k ? aT + idx * N : NULL;
      xs[j] = (uint8_t)(idx < k * k ? idx / k : 0); /* aT[i][j] = XOF(rho, i, j) */
      ys[j] = (uint8_t)(idx < k * k ? idx % k : 0);
    }
    derive_uniform_x4(ps, rho, xs, ys);
  }
  /* EncapsulateTo (kyber.go:103-137) */
  uint8_t g_in[64], kr[64];
  memcpy(g_in, m, 32);
  orc_sha3_256(g_in + 32, ek, orc_mlkem_ek_size(k));
  orc_sha3_512(kr, g_in, 64);
  const uint8_t *seed = kr + 32;
  /* EncryptTo (cpapke.go:137-181) */
  int16_t rh[4 * N], e1[4 * N], u[4 * N], e2[N], v[N], mp[N];
  const int du = du_of(k), dv = dv_of(k);
  for (int i = 0; i < k; i++) orc_kyber_derive_noise(rh + i * N, seed, 32, (uint8_t)i, eta1_of(k));
  for (int i = 0; i < k; i++) {
    orc_kyber_ntt_avx2(rh + i * N);
    orc_kyber_barrett(rh + i * N);
  }
  for (int i = 0; i < k; i++) orc_kyber_derive_noise(e1 + i * N, seed, 32, (uint8_t)(k + i), 2);
  orc_kyber_derive_noise(e2, seed, 32, (uint8_t)(2 * k), 2);
  for (int i = 0; i < k; i++) {
    memset(u + i * N, 0, N * sizeof(int16_t));
    for (int j = 0; j < k; j++) mulhat_acc_avx2(u + i * N, aT + (i * k + j) * N, rh + j * N);
    orc_kyber_barrett(u + i * N);
    orc_kyber_invntt_avx2(u + i * N);
    orc_kyber_add(u + i * N, u + i * N, e1 + i * N);
  }
  memset(v, 0, sizeof v);
  for (int j = 0; j < k; j++) mulhat_acc_avx2(v, th + j * N, rh + j * N);
  orc_kyber_barrett(v);
  orc_kyber_invntt_avx2(v);
  orc_kyber_msg_decompress(mp, m);
  orc_kyber_add(v, v, mp);
  orc_kyber_add(v, v, e2);
  for (int i = 0; i < k; i++) {
    orc_kyber_normalize(u + i * N);
    orc_kyber_compress(ct + i * 32 * du, u + i * N, du);
  }
  orc_kyber_normalize(v);
  orc_kyber_compress(ct + k * 32 * du, v, dv);
  memcpy(ss, kr, 32);
  return 0;
}

typedef struct {
  int k; uint8_t *ct, *ss; const uint8_t *ek; size_t ek_stride; const uint8_t *m; size_t lo, hi; int fails;
} enc_job;
static void *enc_worker(void *arg) {
  enc_job *j = (enc_job *)arg;
  const size_t ctsz = orc_mlkem_ct_size(j->k);
  for (size_t i = j->lo; i < j->hi; i++)
    if (orc_mlkem_encaps_avx2(j->k, j->ct + i * ctsz, j->ss + i * 32, j->ek + i * j->ek_stride, j->m + i * 32)) j->fails++;
  return NULL;
}
int orc_mlkem_encaps_batch_avx2(int k, uint8_t *ct, uint8_t *ss, const uint8_t *ek, size_t ek_stride, const uint8_t *m,
                                size_t n, int nthreads) {
  if (!k_ready) k_tables();
  if (!z_ready) z_tables();
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  enc_job *jobs = (enc_job *)malloc(sizeof(enc_job) * nthreads);
  int fails = 0;
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (enc_job){k, ct, ss, ek, ek_stride, m, n * t / nthreads, n * (t + 1) / nthreads, 0};
    pthread_create(&th[t], NULL, enc_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) {
    pthread_join(th[t], NULL);
    fails += jobs[t].fails;
  }
  free(th);
  free(jobs);
  return fails;
}
