/*
 * oracle/dilithium.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the reference's generic Dilithium / ML-DSA-65 path:
 *   sign/internal/dilithium/{field.go, ntt.go, poly.go, pack.go, params/params.go}
 *   sign/mldsa/mldsa65/internal/{params.go, mat.go, vec.go, sample.go, rounding.go, pack.go, dilithium.go}
 *   sign/mldsa/mldsa65/dilithium.go:56-98 (external-interface framing, unsafeSignInternal)
 * The parameter set is a run-time value (sign/dilithium/gen.go:80-162, NIST=true, tr = 64 B):
 *   ML-DSA-44  K=4 L=4 eta=2 tau=39 gamma1=2^17 gamma2=(q-1)/88 omega=80 c~=32
 *   ML-DSA-65  K=6 L=5 eta=4 tau=49 gamma1=2^19 gamma2=(q-1)/32 omega=55 c~=48
 *   ML-DSA-87  K=8 L=7 eta=2 tau=60 gamma1=2^19 gamma2=(q-1)/32 omega=75 c~=64
 * instead of one generated package per mode.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define N 256
#define Q 8380417u
#define QINV 4236238847u /* -(q^-1) mod 2^32, params.go */
#define ROVER256 41978u
#define D 13
#define KMAX 8
#define LMAX 7
#define TRSIZE 64
#define POLY_T0 416
#define POLY_T1 320

typedef struct {
  int mode, k, l, eta, tau, gamma1_bits, omega, ctilde;
  uint32_t gamma2;
  int nist, tr; /* params.go: NIST, TRSize */
} dparams;
/* modes 44/65/87 = ML-DSA (sign/mldsa/mldsaNN/internal/params.go); modes 2/3/5 = round-3 Dilithium2/3/5
 * (sign/dilithium/modeN/internal/params.go: NIST = false, TRSize = 32, CTildeSize = 32) */
static const dparams MODES[6] = {
    {44, 4, 4, 2, 39, 17, 80, 32, (Q - 1) / 88, 1, 64},
    {65, 6, 5, 4, 49, 19, 55, 48, (Q - 1) / 32, 1, 64},
    {87, 8, 7, 2, 60, 19, 75, 64, (Q - 1) / 32, 1, 64},
    {2, 4, 4, 2, 39, 17, 80, 32, (Q - 1) / 88, 0, 32},
    {3, 6, 5, 4, 49, 19, 55, 32, (Q - 1) / 32, 0, 32},
    {5, 8, 7, 2, 60, 19, 75, 32, (Q - 1) / 32, 0, 32},
};
static const dparams *mode_of(int mode) {
  for (int i = 0; i < 6; i++)
    if (MODES[i].mode == mode) return &MODES[i];
  return NULL;
}
#define P_BETA(p) ((uint32_t)((p)->tau * (p)->eta))
#define P_GAMMA1(p) (1u << (p)->gamma1_bits)
#define P_ALPHA(p) (2 * (p)->gamma2)
#define P_LEQETA(p) ((p)->eta == 2 ? 96 : 128)                 /* PolyLeqEtaSize */
#define P_LEGAMMA1(p) (32 * ((p)->gamma1_bits + 1))            /* PolyLeGamma1Size */
#define P_W1(p) (32 * (23 - (p)->gamma1_bits))                 /* PolyW1Size */
#define P_SK(p) (32 + 32 + (p)->tr + P_LEQETA(p) * ((p)->l + (p)->k) + POLY_T0 * (p)->k)
#define P_PK(p) (32 + POLY_T1 * (p)->k)
#define P_SIG(p) ((p)->l * P_LEGAMMA1(p) + (p)->omega + (p)->k + (p)->ctilde)
size_t orc_mldsa_sk_size(int mode) { return P_SK(mode_of(mode)); }
size_t orc_mldsa_pk_size(int mode) { return P_PK(mode_of(mode)); }
size_t orc_mldsa_sig_size(int mode) { return P_SIG(mode_of(mode)); }

typedef uint32_t poly[N];

/* ---- field.go ---- */
static uint32_t reduce_le2q(uint32_t x) { /* field.go:5-13 */
  uint32_t x1 = x >> 23, x2 = x & 0x7FFFFF;
  return x2 + (x1 << 13) - x1;
}
static uint32_t le2q_modq(uint32_t x) { /* field.go:27-32 */
  x -= Q;
  uint32_t mask = (uint32_t)((int32_t)x >> 31);
  return x + (mask & Q);
}
static uint32_t modq(uint32_t x) { return le2q_modq(reduce_le2q(x)); }
static uint32_t mont_le2q(uint64_t x) { /* field.go:20-24 */
  uint64_t m = (x * QINV) & 0xffffffffu;
  return (uint32_t)((x + m * (uint64_t)Q) >> 32);
}
uint32_t orc_dil_mont_reduce_le2q(uint64_t x) { return mont_le2q(x); }

/* ---- ntt.go: twiddles regenerated from the formulas in ntt.go:3-18 ---- */
static uint32_t ZETAS[N], INVZETAS[N];
static int tw_ready;
static uint32_t powmod(uint32_t b, uint32_t e) {
  uint64_t r = 1, x = b;
  while (e) {
    if (e & 1) r = r * x % Q;
    x = x * x % Q;
    e >>= 1;
  }
  return (uint32_t)r;
}
static unsigned brv8(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 8; i++) r |= ((x >> i) & 1u) << (7 - i);
  return r;
}
static void init_tw(void) {
  if (tw_ready) return;
  uint64_t R = ((uint64_t)1 << 32) % Q;
  uint32_t zinv = powmod(1753, Q - 2);
  for (unsigned i = 0; i < N; i++) {
    ZETAS[i] = (uint32_t)((uint64_t)powmod(1753, brv8(i)) * R % Q);
    INVZETAS[i] = (uint32_t)((uint64_t)powmod(zinv, 256 - brv8(255 - i)) * R % Q);
  }
  tw_ready = 1;
}
const uint32_t *orc_dil_zetas(void) { init_tw(); return ZETAS; }
const uint32_t *orc_dil_inv_zetas(void) { init_tw(); return INVZETAS; }

void orc_dil_ntt(uint32_t p[N]) { /* ntt.go:111-184 */
  init_tw();
  int k = 0;
  for (unsigned l = N / 2; l > 0; l >>= 1)
    for (unsigned off = 0; off < N - l; off += 2 * l) {
      uint64_t zeta = ZETAS[++k];
      for (unsigned j = off; j < off + l; j++) {
        uint32_t t = mont_le2q(zeta * (uint64_t)p[j + l]);
        p[j + l] = p[j] + (2 * Q - t);
        p[j] += t;
      }
    }
}
void orc_dil_invntt(uint32_t p[N]) { /* ntt.go:191-217 */
  init_tw();
  int k = 0;
  for (unsigned l = 1; l < N; l <<= 1)
    for (unsigned off = 0; off < N - l; off += 2 * l) {
      uint64_t zeta = INVZETAS[k++];
      for (unsigned j = off; j < off + l; j++) {
        uint32_t t = p[j];
        p[j] = t + p[j + l];
        t += 256 * Q - p[j + l];
        p[j + l] = mont_le2q(zeta * (uint64_t)t);
      }
    }
  for (unsigned j = 0; j < N; j++) p[j] = mont_le2q((uint64_t)ROVER256 * p[j]);
}

/* ---- poly.go ---- */
void orc_dil_mulhat(uint32_t p[N], const uint32_t a[N], const uint32_t b[N]) { /* poly.go:88 */
  for (int i = 0; i < N; i++) p[i] = mont_le2q((uint64_t)a[i] * b[i]);
}
static void p_add(poly p, const poly a, const poly b) { for (int i = 0; i < N; i++) p[i] = a[i] + b[i]; }
static void p_sub(poly p, const poly a, const poly b) { for (int i = 0; i < N; i++) p[i] = a[i] + (2 * Q - b[i]); }
static void p_reduce_le2q(poly p) { for (int i = 0; i < N; i++) p[i] = reduce_le2q(p[i]); }
static void p_normalize(poly p) { for (int i = 0; i < N; i++) p[i] = modq(p[i]); }
static void p_normalize_le2q(poly p) { for (int i = 0; i < N; i++) p[i] = le2q_modq(p[i]); }
static int p_exceeds(const poly p, uint32_t bound) { /* poly.go:51-71 */
  for (int i = 0; i < N; i++) {
    int32_t x = (int32_t)((Q - 1) / 2) - (int32_t)p[i];
    x ^= (x >> 31);
    x = (int32_t)((Q - 1) / 2) - x;
    if ((uint32_t)x >= bound) return 1;
  }
  return 0;
}
static void power2round(uint32_t a, uint32_t *a0plusq, uint32_t *a1) { /* field.go:35-49 */
  uint32_t a0 = a & ((1u << D) - 1);
  a0 -= (1u << (D - 1)) + 1;
  a0 += (uint32_t)((int32_t)a0 >> 31) & (1u << D);
  a0 -= (1u << (D - 1)) - 1;
  *a0plusq = Q + a0;
  *a1 = (a - a0) >> D;
}
void orc_dil_poly_op(int op, uint32_t *p, const uint32_t *a, const uint32_t *b) {
  switch (op) {
  case 0: p_add(p, a, b); break;
  case 1: p_sub(p, a, b); break;
  case 2: memcpy(p, a, sizeof(poly)); p_reduce_le2q(p); break;
  case 3: memcpy(p, a, sizeof(poly)); p_normalize(p); break;
  case 4: memcpy(p, a, sizeof(poly)); p_normalize_le2q(p); break;
  }
}
int orc_dil_exceeds(const uint32_t *p, uint32_t bound) { return p_exceeds(p, bound); }
void orc_dil_ntt_batch(uint32_t *p, size_t n, int inverse) {
  for (size_t i = 0; i < n; i++) inverse ? orc_dil_invntt(p + i * N) : orc_dil_ntt(p + i * N);
}
void orc_dil_mulhat_batch(uint32_t *p, const uint32_t *a, const uint32_t *b, size_t n) {
  for (size_t i = 0; i < n; i++) orc_dil_mulhat(p + i * N, a + i * N, b + i * N);
}

/* ---- sign/internal/dilithium/pack.go + mldsa65/internal/pack.go ---- */
static void unpack_t1(poly p, const uint8_t *buf) {
  for (int i = 0, j = 0; i < POLY_T1; i += 5, j += 4) {
    p[j] = (buf[i] | ((uint32_t)buf[i + 1] << 8)) & 0x3ff;
    p[j + 1] = ((buf[i + 1] >> 2) | ((uint32_t)buf[i + 2] << 6)) & 0x3ff;
    p[j + 2] = ((buf[i + 2] >> 4) | ((uint32_t)buf[i + 3] << 4)) & 0x3ff;
    p[j + 3] = ((buf[i + 3] >> 6) | ((uint32_t)buf[i + 4] << 2)) & 0x3ff;
  }
}
static void pack_bits(uint8_t *buf, const uint32_t *v, int n, int bits) { /* little-endian bit stream */
  uint64_t acc = 0;
  int have = 0, o = 0;
  for (int i = 0; i < n; i++) {
    acc |= (uint64_t)(v[i] & ((1u << bits) - 1)) << have;
    have += bits;
    while (have >= 8) { buf[o++] = (uint8_t)acc; acc >>= 8; have -= 8; }
  }
}
static void unpack_bits(uint32_t *v, const uint8_t *buf, int n, int bits) {
  uint64_t acc = 0;
  int have = 0, o = 0;
  for (int i = 0; i < n; i++) {
    while (have < bits) { acc |= (uint64_t)buf[o++] << have; have += 8; }
    v[i] = (uint32_t)acc & ((1u << bits) - 1);
    acc >>= bits; have -= bits;
  }
}
static void pack_t1(uint8_t *buf, const poly p) { pack_bits(buf, p, N, 10); } /* pack.go:88-100 */
static void pack_t0(uint8_t *buf, const poly p) {                              /* pack.go:23-54 */
  poly t;
  for (int i = 0; i < N; i++) t[i] = Q + (1u << (D - 1)) - p[i];
  pack_bits(buf, t, N, 13);
}
static void unpack_t0(poly p, const uint8_t *buf) { /* pack.go:56-86 */
  unpack_bits(p, buf, N, 13);
  for (int i = 0; i < N; i++) p[i] = Q + (1u << (D - 1)) - p[i];
}
static void pack_leqeta(const dparams *P, uint8_t *buf, const poly p) { /* internal/pack.go:13-47: 4 or 3 bits */
  poly t;
  for (int i = 0; i < N; i++) t[i] = (uint8_t)(Q + P->eta - p[i]);
  pack_bits(buf, t, N, P->eta == 2 ? 3 : 4);
}
static void unpack_leqeta(const dparams *P, poly p, const uint8_t *buf) { /* internal/pack.go:49-75 */
  unpack_bits(p, buf, N, P->eta == 2 ? 3 : 4);
  for (int i = 0; i < N; i++) p[i] = Q + P->eta - p[i];
}
static void unpack_legamma1(const dparams *P, poly p, const uint8_t *buf) { /* internal/pack.go:146-203: 18 or 20 bits */
  unpack_bits(p, buf, N, P->gamma1_bits + 1);
  for (int i = 0; i < N; i++) {
    uint32_t v = P_GAMMA1(P) - p[i];
    v += (uint32_t)((int32_t)v >> 31) & Q;
    p[i] = v;
  }
}
static void pack_legamma1(const dparams *P, uint8_t *buf, const poly p) { /* internal/pack.go:205-254 */
  poly t;
  for (int i = 0; i < N; i++) {
    uint32_t v = P_GAMMA1(P) - p[i];
    v += (uint32_t)((int32_t)v >> 31) & Q;
    t[i] = v;
  }
  pack_bits(buf, t, N, P->gamma1_bits + 1);
}
/* PolyPackW1 (internal/pack.go:256-271): PackLe16 (4 bits) for gamma1 = 2^19, 6 bits for gamma1 = 2^17 */
static void pack_w1(const dparams *P, uint8_t *buf, const poly p) { pack_bits(buf, p, N, 23 - P->gamma1_bits); }
static void pack_hint(const dparams *P, uint8_t *buf, poly *h) { /* internal/pack.go:77-95 */
  uint8_t off = 0;
  for (int i = 0; i < P->k; i++) {
    for (int j = 0; j < N; j++)
      if (h[i][j] != 0) buf[off++] = (uint8_t)j;
    buf[P->omega + i] = off;
  }
  for (; off < P->omega; off++) buf[off] = 0;
}
static int unpack_hint(const dparams *P, poly *h, const uint8_t *buf) { /* internal/pack.go:113-140 */
  memset(h, 0, sizeof(poly) * P->k);
  uint8_t prev = 0;
  for (int i = 0; i < P->k; i++) {
    uint8_t sop = buf[P->omega + i];
    if (sop < prev || sop > P->omega) return 0;
    for (uint8_t j = prev; j < sop; j++) {
      if (j > prev && buf[j] <= buf[j - 1]) return 0;
      h[i][buf[j]] = 1;
    }
    prev = sop;
  }
  for (uint8_t j = prev; j < P->omega; j++)
    if (buf[j] != 0) return 0;
  return 1;
}

/* ---- sample.go ---- */
void orc_dil_derive_uniform(uint32_t p[N], const uint8_t seed[32], uint16_t nonce) { /* sample.go:92-123 */
  orc_sponge h;
  uint8_t iv[34], buf[168];
  memcpy(iv, seed, 32);
  iv[32] = (uint8_t)nonce; iv[33] = (uint8_t)(nonce >> 8);
  orc_sponge_init(&h, 168, 0x1f);
  orc_sponge_write(&h, iv, 34);
  int i = 0;
  while (i < N) {
    orc_sponge_read(&h, buf, 168);
    for (int j = 0; j < 168 && i < N; j += 3) {
      uint32_t t = (buf[j] | ((uint32_t)buf[j + 1] << 8) | ((uint32_t)buf[j + 2] << 16)) & 0x7fffff;
      if (t < Q) p[i++] = t;
    }
  }
}
static void derive_leqeta(const dparams *P, uint32_t p[N], const uint8_t seed[64], uint16_t nonce) { /* sample.go:129-181 */
  orc_sponge h;
  uint8_t iv[66], buf[136];
  memcpy(iv, seed, 64);
  iv[64] = (uint8_t)nonce; iv[65] = (uint8_t)(nonce >> 8);
  orc_sponge_init(&h, 136, 0x1f);
  orc_sponge_write(&h, iv, 66);
  int i = 0;
  const uint32_t eta = (uint32_t)P->eta;
  while (i < N) {
    orc_sponge_read(&h, buf, 136);
    for (int j = 0; j < 136 && i < N; j++) {
      uint32_t t1 = buf[j] & 15, t2 = buf[j] >> 4;
      if (eta == 2) {
        if (t1 <= 14) { t1 -= ((205 * t1) >> 10) * 5; p[i++] = Q + eta - t1; }
        if (t2 <= 14 && i < N) { t2 -= ((205 * t2) >> 10) * 5; p[i++] = Q + eta - t2; }
      } else {
        if (t1 <= 2 * eta) p[i++] = Q + eta - t1;
        if (t2 <= 2 * eta && i < N) p[i++] = Q + eta - t2;
      }
    }
  }
}
static void derive_legamma1(const dparams *P, uint32_t p[N], const uint8_t seed[64], uint16_t nonce) { /* sample.go:197-209 */
  uint8_t iv[66], buf[640];
  memcpy(iv, seed, 64);
  iv[64] = (uint8_t)nonce; iv[65] = (uint8_t)(nonce >> 8);
  orc_shake256(buf, (size_t)P_LEGAMMA1(P), iv, 66);
  unpack_legamma1(P, p, buf);
}
static void derive_ball(const dparams *P, uint32_t p[N], const uint8_t *seed) { /* sample.go:299-339 */
  orc_sponge h;
  uint8_t buf[136];
  orc_sponge_init(&h, 136, 0x1f);
  orc_sponge_write(&h, seed, (size_t)P->ctilde);
  orc_sponge_read(&h, buf, 136);
  uint64_t signs;
  memcpy(&signs, buf, 8);
  int off = 8;
  memset(p, 0, sizeof(poly));
  for (unsigned i = N - (unsigned)P->tau; i < N; i++) {
    unsigned b;
    for (;;) {
      if (off >= 136) { orc_sponge_read(&h, buf, 136); off = 0; }
      b = buf[off++];
      if (b <= i) break;
    }
    p[i] = p[b];
    p[b] = 1;
    p[b] ^= (uint32_t)((-(signs & 1)) & (1 | (Q - 1)));
    signs >>= 1;
  }
}
/* the ML-DSA-65 instances kept for the sampler-vector tests */
void orc_dil_derive_leqeta(uint32_t p[N], const uint8_t seed[64], uint16_t nonce) { derive_leqeta(mode_of(65), p, seed, nonce); }
void orc_dil_derive_legamma1(uint32_t p[N], const uint8_t seed[64], uint16_t nonce) { derive_legamma1(mode_of(65), p, seed, nonce); }
void orc_dil_derive_ball(uint32_t p[N], const uint8_t seed[48]) { derive_ball(mode_of(65), p, seed); }
void orc_mldsa_derive_leqeta(int mode, uint32_t p[N], const uint8_t seed[64], uint16_t nonce) { derive_leqeta(mode_of(mode), p, seed, nonce); }
void orc_mldsa_derive_legamma1(int mode, uint32_t p[N], const uint8_t seed[64], uint16_t nonce) { derive_legamma1(mode_of(mode), p, seed, nonce); }
void orc_mldsa_derive_ball(int mode, uint32_t p[N], const uint8_t *seed) { derive_ball(mode_of(mode), p, seed); }
void orc_dil_power2round(const uint32_t *p, uint32_t *p0plusq, uint32_t *p1) { /* poly.go:77-84 */
  for (int i = 0; i < N; i++) power2round(p[i], &p0plusq[i], &p1[i]);
}
void orc_dil_pack_le16(uint8_t buf[128], const uint32_t *p) { /* pack.go:102-108 */
  for (int i = 0; i < 128; i++) buf[i] = (uint8_t)((uint8_t)p[2 * i] | (uint8_t)(p[2 * i + 1] << 4));
}

/* ---- rounding.go ---- */
static void decompose(const dparams *P, uint32_t a, uint32_t *a0plusq, uint32_t *a1o) { /* rounding.go:13-43 */
  uint32_t a1 = (a + 127) >> 7;
  if (P_ALPHA(P) == 523776) {
    a1 = (a1 * 1025 + (1u << 21)) >> 22;
    a1 &= 15;
  } else { /* alpha = 190464 */
    a1 = (a1 * 11275 + (1u << 23)) >> 24;
    a1 ^= (uint32_t)((int32_t)(43 - a1) >> 31) & a1;
  }
  uint32_t a0 = a - a1 * P_ALPHA(P);
  a0 += (uint32_t)((int32_t)(a0 - (Q - 1) / 2) >> 31) & Q;
  *a0plusq = a0;
  *a1o = a1;
}
static uint32_t make_hint(const dparams *P, uint32_t z0, uint32_t r1) { /* rounding.go:56-67 */
  if (z0 <= P->gamma2 || z0 > Q - P->gamma2 || (z0 == Q - P->gamma2 && r1 == 0)) return 0;
  return 1;
}
void orc_dil_decompose(const uint32_t *p, uint32_t *p0, uint32_t *p1) {
  for (int i = 0; i < N; i++) decompose(mode_of(65), p[i], &p0[i], &p1[i]);
}

/* ---- private key (dilithium.go:56-72, Unpack :142-163) ---- */
typedef struct {
  uint8_t rho[32], key[32], tr[TRSIZE];
  poly s1[LMAX], s2[KMAX], t0[KMAX];
  poly A[KMAX][LMAX], s1h[LMAX], s2h[KMAX], t0h[KMAX];
} privkey;

static void mat_derive(const dparams *P, poly A[KMAX][LMAX], const uint8_t rho[32]) { /* mat.go:15-23 */
  for (int i = 0; i < P->k; i++)
    for (int j = 0; j < P->l; j++) orc_dil_derive_uniform(A[i][j], rho, (uint16_t)((i << 8) + j));
}
static void dot_hat(const dparams *P, poly p, poly *a, poly *b) { /* mat.go:52-59 */
  poly t;
  memset(p, 0, sizeof(poly));
  for (int i = 0; i < P->l; i++) {
    orc_dil_mulhat(t, a[i], b[i]);
    p_add(p, t, p);
  }
}
static void sk_cache(const dparams *P, privkey *sk) {
  mat_derive(P, sk->A, sk->rho);
  for (int i = 0; i < P->k; i++) { memcpy(sk->t0h[i], sk->t0[i], sizeof(poly)); orc_dil_ntt(sk->t0h[i]); }
  for (int i = 0; i < P->l; i++) { memcpy(sk->s1h[i], sk->s1[i], sizeof(poly)); orc_dil_ntt(sk->s1h[i]); }
  for (int i = 0; i < P->k; i++) { memcpy(sk->s2h[i], sk->s2[i], sizeof(poly)); orc_dil_ntt(sk->s2h[i]); }
}
static void sk_unpack(const dparams *P, privkey *sk, const uint8_t *buf) {
  memcpy(sk->rho, buf, 32);
  memcpy(sk->key, buf + 32, 32);
  memcpy(sk->tr, buf + 64, (size_t)P->tr);
  const uint8_t *p = buf + 64 + P->tr;
  for (int i = 0; i < P->l; i++, p += P_LEQETA(P)) unpack_leqeta(P, sk->s1[i], p);
  for (int i = 0; i < P->k; i++, p += P_LEQETA(P)) unpack_leqeta(P, sk->s2[i], p);
  for (int i = 0; i < P->k; i++, p += POLY_T0) unpack_t0(sk->t0[i], p);
  sk_cache(P, sk);
}
static void sk_pack(const dparams *P, const privkey *sk, uint8_t *buf) { /* dilithium.go:129-139 */
  memcpy(buf, sk->rho, 32);
  memcpy(buf + 32, sk->key, 32);
  memcpy(buf + 64, sk->tr, (size_t)P->tr);
  uint8_t *p = buf + 64 + P->tr;
  for (int i = 0; i < P->l; i++, p += P_LEQETA(P)) pack_leqeta(P, p, sk->s1[i]);
  for (int i = 0; i < P->k; i++, p += P_LEQETA(P)) pack_leqeta(P, p, sk->s2[i]);
  for (int i = 0; i < P->k; i++, p += POLY_T0) pack_t0(p, sk->t0[i]);
}

/* NewKeyFromSeed, dilithium.go:181-241 (+ computeT0andT1 :253-267) */
void orc_mldsa_keygen(int mode, uint8_t *pk, uint8_t *skb, const uint8_t seed[32]) {
  const dparams *P = mode_of(mode);
  privkey *sk = (privkey *)malloc(sizeof(privkey));
  uint8_t in[34], eseed[128];
  memcpy(in, seed, 32);
  in[32] = (uint8_t)P->k; in[33] = (uint8_t)P->l;
  orc_shake256(eseed, 128, in, P->nist ? 34 : 32); /* dilithium.go:191-193: K, L only when NIST */
  memcpy(sk->rho, eseed, 32);
  const uint8_t *sseed = eseed + 32;
  memcpy(sk->key, eseed + 96, 32);
  for (int i = 0; i < P->l; i++) derive_leqeta(P, sk->s1[i], sseed, (uint16_t)i);
  for (int i = 0; i < P->k; i++) derive_leqeta(P, sk->s2[i], sseed, (uint16_t)(i + P->l));
  mat_derive(P, sk->A, sk->rho);
  for (int i = 0; i < P->l; i++) { memcpy(sk->s1h[i], sk->s1[i], sizeof(poly)); orc_dil_ntt(sk->s1h[i]); }
  poly t1[KMAX];
  for (int i = 0; i < P->k; i++) {
    poly t;
    dot_hat(P, t, sk->A[i], sk->s1h);
    p_reduce_le2q(t);
    orc_dil_invntt(t);
    p_add(t, t, sk->s2[i]);
    p_normalize(t);
    for (int j = 0; j < N; j++) power2round(t[j], &sk->t0[i][j], &t1[i][j]);
  }
  memcpy(pk, sk->rho, 32);
  for (int i = 0; i < P->k; i++) pack_t1(pk + 32 + POLY_T1 * i, t1[i]);
  orc_shake256(sk->tr, (size_t)P->tr, pk, (size_t)P_PK(P));
  sk_pack(P, sk, skb);
  free(sk);
}

/* ML-DSA.Sign_internal: internal/dilithium.go:340-470.  msg = M' (already framed by the caller).
 * Returns the number of attempts, or -1 after 576. */
static int sign_internal(const dparams *P, const privkey *sk, const uint8_t *msg, size_t msglen, const uint8_t rnd[32],
                         uint8_t *sig) {
  uint8_t mu[64], rhop[64], w1p[192 * KMAX], ctilde[64];
  const int K = P->k, L = P->l;
  orc_sponge h;
  orc_sponge_init(&h, 136, 0x1f);
  orc_sponge_write(&h, sk->tr, (size_t)P->tr);
  orc_sponge_write(&h, msg, msglen);
  orc_sponge_read(&h, mu, 64);
  orc_sponge_init(&h, 136, 0x1f);
  orc_sponge_write(&h, sk->key, 32);
  if (P->nist) orc_sponge_write(&h, rnd, 32); /* dilithium.go:360-362 */
  orc_sponge_write(&h, mu, 64);
  orc_sponge_read(&h, rhop, 64);

  poly *y = malloc(sizeof(poly) * LMAX), *yh = malloc(sizeof(poly) * LMAX), *z = malloc(sizeof(poly) * LMAX);
  poly *w = malloc(sizeof(poly) * KMAX), *w0 = malloc(sizeof(poly) * KMAX), *w1 = malloc(sizeof(poly) * KMAX);
  poly *w0mcs2 = malloc(sizeof(poly) * KMAX), *ct0 = malloc(sizeof(poly) * KMAX), *hint = malloc(sizeof(poly) * KMAX);
  poly ch;
  uint16_t ynonce = 0;
  int attempt = 0, ok = 0;
  while (!ok) {
    attempt++;
    if (attempt >= 576) { attempt = -1; break; }
    for (int i = 0; i < L; i++) derive_legamma1(P, y[i], rhop, (uint16_t)(ynonce + i));
    ynonce = (uint16_t)(ynonce + L);
    for (int i = 0; i < L; i++) { memcpy(yh[i], y[i], sizeof(poly)); orc_dil_ntt(yh[i]); }
    for (int i = 0; i < K; i++) {
      dot_hat(P, w[i], ((privkey *)sk)->A[i], yh);
      p_reduce_le2q(w[i]);
      orc_dil_invntt(w[i]);
      p_normalize_le2q(w[i]);
      for (int j = 0; j < N; j++) decompose(P, w[i][j], &w0[i][j], &w1[i][j]);
      pack_w1(P, w1p + P_W1(P) * i, w1[i]);
    }
    orc_sponge_init(&h, 136, 0x1f);
    orc_sponge_write(&h, mu, 64);
    orc_sponge_write(&h, w1p, (size_t)(P_W1(P) * K));
    orc_sponge_read(&h, ctilde, (size_t)P->ctilde);
    derive_ball(P, ch, ctilde);
    orc_dil_ntt(ch);
    int rej = 0;
    for (int i = 0; i < K; i++) {
      orc_dil_mulhat(w0mcs2[i], ch, sk->s2h[i]);
      orc_dil_invntt(w0mcs2[i]);
      p_sub(w0mcs2[i], w0[i], w0mcs2[i]);
      p_normalize(w0mcs2[i]);
      rej |= p_exceeds(w0mcs2[i], P->gamma2 - P_BETA(P));
    }
    if (rej) continue;
    for (int i = 0; i < L; i++) {
      orc_dil_mulhat(z[i], ch, sk->s1h[i]);
      orc_dil_invntt(z[i]);
      p_add(z[i], z[i], y[i]);
      p_normalize(z[i]);
      rej |= p_exceeds(z[i], P_GAMMA1(P) - P_BETA(P));
    }
    if (rej) continue;
    for (int i = 0; i < K; i++) {
      orc_dil_mulhat(ct0[i], ch, sk->t0h[i]);
      orc_dil_invntt(ct0[i]);
      p_normalize_le2q(ct0[i]);
      rej |= p_exceeds(ct0[i], P->gamma2);
    }
    if (rej) continue;
    uint32_t pop = 0;
    for (int i = 0; i < K; i++) {
      poly s;
      p_add(s, w0mcs2[i], ct0[i]);
      p_normalize_le2q(s);
      for (int j = 0; j < N; j++) { hint[i][j] = make_hint(P, s[j], w1[i][j]); pop += hint[i][j]; }
    }
    if (pop > (uint32_t)P->omega) continue;
    ok = 1;
  }
  if (ok) {
    memcpy(sig, ctilde, (size_t)P->ctilde);
    for (int i = 0; i < L; i++) pack_legamma1(P, sig + P->ctilde + P_LEGAMMA1(P) * i, z[i]);
    pack_hint(P, sig + P->ctilde + L * P_LEGAMMA1(P), hint);
  }
  free(y); free(yh); free(z); free(w); free(w0); free(w1); free(w0mcs2); free(ct0); free(hint);
  return attempt;
}

/* sk: packed.  internal != 0: ML-DSA.Sign_internal on msg as given (ACVP internal interface,
 * mldsa65/dilithium.go:87-98); else the external framing 0x00 || len(ctx) || ctx || msg
 * (mldsa65/dilithium.go:56-84).  rnd: 32 bytes (all zero = deterministic).  Returns attempts or -1. */
int orc_mldsa_sign(int mode, uint8_t *sig, const uint8_t *skb, const uint8_t *msg, size_t msglen, const uint8_t *ctx,
                   size_t ctxlen, const uint8_t rnd[32], int internal) {
  const dparams *P = mode_of(mode);
  privkey *sk = (privkey *)malloc(sizeof(privkey));
  sk_unpack(P, sk, skb);
  int rc;
  if (internal || !P->nist) { /* round 3 signs the message as given (sign/dilithium/mode3/dilithium.go:54-64) */
    rc = sign_internal(P, sk, msg, msglen, rnd, sig);
  } else {
    uint8_t *m = (uint8_t *)malloc(2 + ctxlen + msglen);
    m[0] = 0; m[1] = (uint8_t)ctxlen;
    if (ctxlen) memcpy(m + 2, ctx, ctxlen);
    memcpy(m + 2 + ctxlen, msg, msglen);
    rc = sign_internal(P, sk, m, 2 + ctxlen + msglen, rnd, sig);
    free(m);
  }
  free(sk);
  return rc;
}

/* Verify (internal/dilithium.go:273-332); same message conventions as sign. returns 1 = valid */
int orc_mldsa_verify(int mode, const uint8_t *pkb, const uint8_t *msg, size_t msglen, const uint8_t *ctx, size_t ctxlen,
                     const uint8_t *sig, size_t siglen, int internal) {
  const dparams *P = mode_of(mode);
  const int K = P->k, L = P->l;
  if (siglen != (size_t)P_SIG(P)) return 0;
  poly *z = malloc(sizeof(poly) * LMAX), *hint = malloc(sizeof(poly) * KMAX), *t1 = malloc(sizeof(poly) * KMAX);
  poly(*A)[LMAX] = malloc(sizeof(poly) * KMAX * LMAX);
  int ok = 0;
  for (int i = 0; i < L; i++) unpack_legamma1(P, z[i], sig + P->ctilde + P_LEGAMMA1(P) * i);
  for (int i = 0; i < L; i++)
    if (p_exceeds(z[i], P_GAMMA1(P) - P_BETA(P))) goto done;
  if (!unpack_hint(P, hint, sig + P->ctilde + L * P_LEGAMMA1(P))) goto done;
  {
    uint8_t tr[TRSIZE], mu[64], w1p[192 * KMAX], cp[64];
    orc_shake256(tr, (size_t)P->tr, pkb, (size_t)P_PK(P));
    orc_sponge h;
    orc_sponge_init(&h, 136, 0x1f);
    orc_sponge_write(&h, tr, (size_t)P->tr);
    if (!internal && P->nist) {
      uint8_t pre[2] = {0, (uint8_t)ctxlen};
      orc_sponge_write(&h, pre, 2);
      orc_sponge_write(&h, ctx, ctxlen);
    }
    orc_sponge_write(&h, msg, msglen);
    orc_sponge_read(&h, mu, 64);
    mat_derive(P, A, pkb);
    for (int i = 0; i < K; i++) unpack_t1(t1[i], pkb + 32 + POLY_T1 * i);
    for (int i = 0; i < L; i++) orc_dil_ntt(z[i]);
    poly ch;
    derive_ball(P, ch, sig);
    orc_dil_ntt(ch);
    for (int i = 0; i < K; i++) {
      poly az, t, q0, w1;
      dot_hat(P, az, A[i], z);
      for (int j = 0; j < N; j++) t[j] = t1[i][j] << D;
      orc_dil_ntt(t);
      orc_dil_mulhat(t, t, ch);
      p_sub(t, az, t);
      p_reduce_le2q(t);
      orc_dil_invntt(t);
      p_normalize_le2q(t);
      for (int j = 0; j < N; j++) decompose(P, t[j], &q0[j], &w1[j]); /* PolyUseHint, rounding.go:98-135 */
      for (int j = 0; j < N; j++) {
        if (!hint[i][j]) continue;
        if (P->gamma2 == 261888) {
          w1[j] = (q0[j] > Q) ? ((w1[j] + 1) & 15) : ((w1[j] - 1) & 15);
        } else if (q0[j] > Q) {
          w1[j] = (w1[j] == 43) ? 0 : w1[j] + 1;
        } else {
          w1[j] = (w1[j] == 0) ? 43 : w1[j] - 1;
        }
      }
      pack_w1(P, w1p + P_W1(P) * i, w1);
    }
    orc_sponge_init(&h, 136, 0x1f);
    orc_sponge_write(&h, mu, 64);
    orc_sponge_write(&h, w1p, (size_t)(P_W1(P) * K));
    orc_sponge_read(&h, cp, (size_t)P->ctilde);
    ok = memcmp(cp, sig, (size_t)P->ctilde) == 0;
  }
done:
  free(z); free(hint); free(t1); free(A);
  return ok;
}

#include <pthread.h>
typedef struct {
  const dparams *P; uint8_t *sig; const uint8_t *sk; size_t sk_stride; const uint8_t *msgs; const uint64_t *off;
  const uint8_t *rnd; size_t lo, hi; int fails; long attempts;
} sign_job;
static void *sign_worker(void *arg) {
  sign_job *j = (sign_job *)arg;
  static const uint8_t zero[32] = {0};
  privkey *sk = (privkey *)malloc(sizeof(privkey));
  const uint8_t *last = NULL;
  for (size_t i = j->lo; i < j->hi; i++) {
    const uint8_t *skb = j->sk + i * j->sk_stride;
    if (skb != last) { sk_unpack(j->P, sk, skb); last = skb; } /* per-op sk is re-expanded, like the GPU path */
    const uint8_t *m = j->msgs + j->off[i];
    size_t mlen = (size_t)(j->off[i + 1] - j->off[i]);
    uint8_t *fm = (uint8_t *)malloc(2 + mlen);
    fm[0] = 0; fm[1] = 0;
    memcpy(fm + 2, m, mlen);
    const int pre = j->P->nist ? 0 : 2; /* round 3: no framing */
    int a = sign_internal(j->P, sk, fm + pre, 2 + mlen - pre, j->rnd ? j->rnd + 32 * i : zero,
                          j->sig + i * (size_t)P_SIG(j->P));
    free(fm);
    if (a < 0) j->fails++; else j->attempts += a;
  }
  free(sk);
  return NULL;
}
/* batched external-interface signing with empty context; msgs concatenated, off[n+1] offsets.
 * returns total attempts (>0) or -1 on failure */
int orc_mldsa_sign_batch(int mode, uint8_t *sig, const uint8_t *sk, size_t sk_stride, const uint8_t *msgs,
                         const uint64_t *off, const uint8_t *rnd, size_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  sign_job *jobs = (sign_job *)malloc(sizeof(sign_job) * nthreads);
  long attempts = 0;
  int fails = 0;
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (sign_job){mode_of(mode), sig, sk, sk_stride, msgs, off, rnd, n * t / nthreads, n * (t + 1) / nthreads, 0, 0};
    pthread_create(&th[t], NULL, sign_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); fails += jobs[t].fails; attempts += jobs[t].attempts; }
  free(th); free(jobs);
  return fails ? -1 : (int)attempts;
}

/* ML-DSA-65 entry points kept under their original names */
void orc_mldsa65_keygen(uint8_t pk[1952], uint8_t sk[4032], const uint8_t seed[32]) { orc_mldsa_keygen(65, pk, sk, seed); }
int orc_mldsa65_sign(uint8_t sig[3309], const uint8_t *sk, const uint8_t *msg, size_t msglen, const uint8_t *ctx,
                     size_t ctxlen, const uint8_t rnd[32], int internal) {
  return orc_mldsa_sign(65, sig, sk, msg, msglen, ctx, ctxlen, rnd, internal);
}
int orc_mldsa65_verify(const uint8_t pk[1952], const uint8_t *msg, size_t msglen, const uint8_t *ctx, size_t ctxlen,
                       const uint8_t *sig, size_t siglen, int internal) {
  return orc_mldsa_verify(65, pk, msg, msglen, ctx, ctxlen, sig, siglen, internal);
}
int orc_mldsa65_sign_batch(uint8_t *sig, const uint8_t *sk, size_t sk_stride, const uint8_t *msgs, const uint64_t *off,
                           const uint8_t *rnd, size_t n, int nthreads) {
  return orc_mldsa_sign_batch(65, sig, sk, sk_stride, msgs, off, rnd, n, nthreads);
}
