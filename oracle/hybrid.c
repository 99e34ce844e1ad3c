/* oracle/hybrid.c -- TEST INFRASTRUCTURE ONLY (see oracle.h): CPU restatement of the callers on the wire side of
 * the ML-KEM core:
 *   dh/x25519/key.go:1-56, curve.go:40-75      X25519 KeyGen / Shared (RFC 7748 function, low-order check)
 *   kem/xwing/xwing.go:47-66,108-130,209-281   X-Wing: key derivation, Encapsulate, Decapsulate, combiner
 *   kem/hybrid/hybrid.go:197-283, xkem.go      X25519MLKEM768, Kyber768-X25519, Kyber512-X25519
 * The reference computes KeyGen with a Joye ladder over a precomputed table and Shared with a Montgomery ladder
 * (curve.go:7-75); both return the affine u-coordinate of [clamp(k)]P in canonical form, which is what the plain
 * RFC 7748 ladder below returns.  Pinned by the RFC 7748 / Wycheproof vectors the reference's own tests use
 * (dh/x25519/testdata) and by the X-Wing draft vectors hash of kem/xwing/xwing_test.go:40-83.
 */
#include <string.h>

#include "oracle.h"

/* ---------------------------------------------------------------- GF(2^255 - 19), five 51-bit limbs */
typedef uint64_t fe[5];
typedef unsigned __int128 u128;
#define M51 ((1ull << 51) - 1)

static void fe_frombytes(fe h, const uint8_t s[32]) {
  uint64_t w[4];
  memcpy(w, s, 32);
  h[0] = w[0] & M51;
  h[1] = ((w[0] >> 51) | (w[1] << 13)) & M51;
  h[2] = ((w[1] >> 38) | (w[2] << 26)) & M51;
  h[3] = ((w[2] >> 25) | (w[3] << 39)) & M51;
  h[4] = (w[3] >> 12) & M51; /* drops bit 255 */
}
static void fe_carry(fe h) {
  for (int r = 0; r < 2; r++) {
    for (int i = 0; i < 4; i++) { h[i + 1] += h[i] >> 51; h[i] &= M51; }
    h[0] += 19 * (h[4] >> 51); h[4] &= M51;
  }
}
static void fe_tobytes(uint8_t s[32], const fe f) { /* canonical encoding */
  fe h;
  memcpy(h, f, sizeof(fe));
  fe_carry(h);
  /* h < 2^255 + small; subtract p if h >= p: compute h + 19 and look at bit 255 */
  uint64_t q = (h[0] + 19) >> 51;
  for (int i = 1; i < 5; i++) q = (h[i] + q) >> 51;
  h[0] += 19 * q;
  for (int i = 0; i < 4; i++) { h[i + 1] += h[i] >> 51; h[i] &= M51; }
  h[4] &= M51;
  uint64_t w[4] = {h[0] | (h[1] << 51), (h[1] >> 13) | (h[2] << 38), (h[2] >> 26) | (h[3] << 25), (h[3] >> 39) | (h[4] << 12)};
  memcpy(s, w, 32);
}
static void fe_add(fe h, const fe f, const fe g) { for (int i = 0; i < 5; i++) h[i] = f[i] + g[i]; }
static void fe_sub(fe h, const fe f, const fe g) { /* f - g + 2p keeps limbs positive */
  h[0] = f[0] + 0xfffffffffffdaull - g[0];
  for (int i = 1; i < 5; i++) h[i] = f[i] + 0xffffffffffffeull - g[i];
}
static void fe_mul(fe h, const fe f, const fe g) {
  u128 t[5];
  uint64_t g19[5];
  for (int i = 0; i < 5; i++) g19[i] = 19 * g[i];
  for (int k = 0; k < 5; k++) {
    t[k] = 0;
    for (int i = 0; i < 5; i++) {
      int j = k - i;
      t[k] += (u128)f[i] * (j >= 0 ? g[j] : g19[j + 5]);
    }
  }
  u128 c = 0;
  for (int k = 0; k < 5; k++) { t[k] += c; h[k] = (uint64_t)t[k] & M51; c = t[k] >> 51; }
  c = c * 19 + h[0];
  h[0] = (uint64_t)c & M51;
  h[1] += (uint64_t)(c >> 51);
}
static void fe_mul_small(fe h, const fe f, uint64_t s) {
  u128 c = 0;
  for (int k = 0; k < 5; k++) { c += (u128)f[k] * s; h[k] = (uint64_t)c & M51; c >>= 51; }
  h[0] += 19 * (uint64_t)c;
  h[1] += h[0] >> 51; h[0] &= M51;
}
static void fe_pow2k(fe h, const fe f, int k) { fe_mul(h, f, f); for (int i = 1; i < k; i++) fe_mul(h, h, h); }
static void fe_invert(fe out, const fe z) { /* z^(p-2) */
  fe z2, z9, z11, z5, z10, z20, z50, z100, t;
  fe_pow2k(z2, z, 1);
  fe_pow2k(t, z2, 2); fe_mul(z9, t, z);
  fe_mul(z11, z9, z2);
  fe_pow2k(t, z11, 1); fe_mul(z5, t, z9);     /* 2^5 - 1 */
  fe_pow2k(t, z5, 5); fe_mul(z10, t, z5);     /* 2^10 - 1 */
  fe_pow2k(t, z10, 10); fe_mul(z20, t, z10);  /* 2^20 - 1 */
  fe_pow2k(t, z20, 20); fe_mul(t, t, z20);    /* 2^40 - 1 */
  fe_pow2k(t, t, 10); fe_mul(z50, t, z10);    /* 2^50 - 1 */
  fe_pow2k(t, z50, 50); fe_mul(z100, t, z50); /* 2^100 - 1 */
  fe_pow2k(t, z100, 100); fe_mul(t, t, z100); /* 2^200 - 1 */
  fe_pow2k(t, t, 50); fe_mul(t, t, z50);      /* 2^250 - 1 */
  fe_pow2k(t, t, 5); fe_mul(out, t, z11);     /* 2^255 - 21 */
}
static void fe_cswap(fe a, fe b, uint64_t bit) {
  uint64_t m = 0 - bit;
  for (int i = 0; i < 5; i++) { uint64_t x = m & (a[i] ^ b[i]); a[i] ^= x; b[i] ^= x; }
}

/* out = u([clamp(scalar)] P); point == NULL: P = base point (KeyGen, key.go:44-46), else Shared (key.go:48-56).
 * Returns 1 unless `point` (bit 255 cleared, reduced mod p) is one of the five low-order points of curve.go:89-125. */
int orc_x25519(uint8_t out[32], const uint8_t scalar[32], const uint8_t *point) {
  static const uint8_t base[32] = {9};
  uint8_t k[32], canon[32];
  memcpy(k, scalar, 32);
  k[0] &= 248; k[31] = (uint8_t)((k[31] & 127) | 64);
  fe x1, x2 = {1}, z2 = {0}, x3, z3 = {1}, a, aa, b, bb, e, c, d, da, cb, t;
  fe_frombytes(x1, point ? point : base);
  memcpy(x3, x1, sizeof(fe));
  uint64_t swap = 0;
  for (int s = 254; s >= 0; s--) {
    uint64_t bit = (k[s >> 3] >> (s & 7)) & 1;
    swap ^= bit;
    fe_cswap(x2, x3, swap); fe_cswap(z2, z3, swap);
    swap = bit;
    fe_add(a, x2, z2); fe_mul(aa, a, a);
    fe_sub(b, x2, z2); fe_mul(bb, b, b);
    fe_sub(e, aa, bb);
    fe_add(c, x3, z3); fe_sub(d, x3, z3);
    fe_mul(da, d, a); fe_mul(cb, c, b);
    fe_add(t, da, cb); fe_mul(x3, t, t);
    fe_sub(t, da, cb); fe_mul(t, t, t); fe_mul(z3, x1, t);
    fe_mul(x2, aa, bb);
    fe_mul_small(t, e, 121665); fe_add(t, t, aa); fe_mul(z2, e, t);
  }
  fe_cswap(x2, x3, swap); fe_cswap(z2, z3, swap);
  fe_invert(z2, z2);
  fe_mul(x2, x2, z2);
  fe_tobytes(out, x2);
  if (!point) return 1;
  static const uint8_t low[5][32] = {
      {0}, {1},
      {0xe0, 0xeb, 0x7a, 0x7c, 0x3b, 0x41, 0xb8, 0xae, 0x16, 0x56, 0xe3, 0xfa, 0xf1, 0x9f, 0xc4, 0x6a,
       0xda, 0x09, 0x8d, 0xeb, 0x9c, 0x32, 0xb1, 0xfd, 0x86, 0x62, 0x05, 0x16, 0x5f, 0x49, 0xb8, 0x00},
      {0x5f, 0x9c, 0x95, 0xbc, 0xa3, 0x50, 0x8c, 0x24, 0xb1, 0xd0, 0xb1, 0x55, 0x9c, 0x83, 0xef, 0x5b,
       0x04, 0x44, 0x5c, 0xc4, 0x58, 0x1c, 0x8e, 0x86, 0xd8, 0x22, 0x4e, 0xdd, 0xd0, 0x9f, 0x11, 0x57},
      {0xec, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff,
       0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0xff, 0x7f}};
  fe_tobytes(canon, x1); /* bit 255 cleared and reduced mod p, as isValidPubKey does (key.go:25-32) */
  int ok = 1;
  for (int i = 0; i < 5; i++) ok &= memcmp(canon, low[i], 32) != 0;
  return ok;
}

/* ---------------------------------------------------------------- X-Wing (kem/xwing/xwing.go) */
static void xwing_combiner(uint8_t ss[32], const uint8_t ssm[32], const uint8_t ssx[32], const uint8_t ctx[32],
                           const uint8_t pkx[32]) { /* xwing.go:47-66 */
  uint8_t in[134];
  memcpy(in, ssm, 32); memcpy(in + 32, ssx, 32); memcpy(in + 64, ctx, 32); memcpy(in + 96, pkx, 32);
  memcpy(in + 128, "\\.//^\\", 6);
  orc_sha3_256(ss, in, 134);
}
/* deriveKeyPair (xwing.go:108-130): sk = seed; expanded = SHAKE256(seed, 96) = ML-KEM-768 seed (64) || X25519 sk (32) */
static void xwing_expand(const uint8_t seed[32], uint8_t *ek, uint8_t *dk, uint8_t skx[32], uint8_t pkx[32]) {
  uint8_t ex[96];
  orc_shake256(ex, 96, seed, 32);
  orc_mlkem_keygen(3, ek, dk, ex);
  memcpy(skx, ex + 64, 32);
  orc_x25519(pkx, skx, NULL);
}
void orc_xwing_keygen(uint8_t pk[1216], const uint8_t seed[32]) {
  uint8_t dk[2400], skx[32];
  xwing_expand(seed, pk, dk, skx, pk + 1184);
}
/* EncapsulateTo (xwing.go:209-247); eseed = ML-KEM m (32) || ephemeral X25519 secret (32).  rc != 0: kem.ErrPubKey */
int orc_xwing_encaps(uint8_t ct[1120], uint8_t ss[32], const uint8_t pk[1216], const uint8_t eseed[64]) {
  uint8_t ssm[32], ssx[32];
  orc_x25519(ct + 1088, eseed + 32, NULL);
  orc_x25519(ssx, eseed + 32, pk + 1184);
  if (orc_mlkem_encaps(3, ct, ssm, pk, eseed)) return 1;
  xwing_combiner(ss, ssm, ssx, ct + 1088, pk + 1184);
  return 0;
}
void orc_xwing_decaps(uint8_t ss[32], const uint8_t sk[32], const uint8_t ct[1120]) { /* xwing.go:249-272 */
  uint8_t ek[1184], dk[2400], skx[32], pkx[32], ssm[32], ssx[32];
  xwing_expand(sk, ek, dk, skx, pkx);
  orc_mlkem_decaps(3, ssm, dk, ct);
  orc_x25519(ssx, skx, ct + 1088);
  xwing_combiner(ss, ssm, ssx, ct + 1088, pkx);
}

/* ---------------------------------------------------------------- kem/hybrid (hybrid.go, xkem.go)
 * id 0: X25519MLKEM768 (first = ML-KEM-768, second = X25519)      hybrid.go:58-62
 * id 1: Kyber768-X25519 (first = X25519, second = Kyber768)        hybrid.go:40-44
 * id 2: Kyber512-X25519 (first = X25519, second = Kyber512)        hybrid.go:34-38 */
typedef struct { int k, mlkem, x_first; } hyb;
static const hyb HYB[3] = {{3, 1, 0}, {3, 0, 1}, {2, 0, 1}};
size_t orc_hybrid_pk_size(int id) { return orc_mlkem_ek_size(HYB[id].k) + 32; }
size_t orc_hybrid_sk_size(int id) { return orc_mlkem_dk_size(HYB[id].k) + 32; }
size_t orc_hybrid_ct_size(int id) { return orc_mlkem_ct_size(HYB[id].k) + 32; }

/* xScheme.DeriveKeyPair (xkem.go:118-129): sk = SHAKE256(seed, 32), pk = KeyGen(sk) */
static void xkem_derive(uint8_t pk[32], uint8_t sk[32], const uint8_t seed[32]) {
  orc_shake256(sk, 32, seed, 32);
  orc_x25519(pk, sk, NULL);
}
/* scheme.DeriveKeyPair (hybrid.go:197-212): SHAKE256(seed) -> first.SeedSize || second.SeedSize bytes */
void orc_hybrid_keygen(int id, uint8_t *pk, uint8_t *sk, const uint8_t seed[64]) {
  const hyb *H = &HYB[id];
  const size_t eksz = orc_mlkem_ek_size(H->k), dksz = orc_mlkem_dk_size(H->k);
  uint8_t ex[96];
  orc_shake256(ex, 96, seed, 64);
  const uint8_t *sx = H->x_first ? ex : ex + 64, *sm = H->x_first ? ex + 32 : ex;
  uint8_t *pkx = H->x_first ? pk : pk + eksz, *pkm = H->x_first ? pk + 32 : pk;
  uint8_t *skx = H->x_first ? sk : sk + dksz, *skm = H->x_first ? sk + 32 : sk;
  xkem_derive(pkx, skx, sx);
  if (H->mlkem) orc_mlkem_keygen(H->k, pkm, skm, sm); else orc_kyber_kem_keygen(H->k, pkm, skm, sm);
}
/* EncapsulateDeterministically (hybrid.go:233-261, xkem.go:166-183).  rc: 0 ok, 1 kem.ErrPubKey */
int orc_hybrid_encaps(int id, uint8_t *ct, uint8_t *ss, const uint8_t *pk, const uint8_t seed[32]) {
  const hyb *H = &HYB[id];
  const size_t eksz = orc_mlkem_ek_size(H->k), ctsz = orc_mlkem_ct_size(H->k);
  uint8_t ex[64], esk[32];
  orc_shake256(ex, 64, seed, 32);
  const uint8_t *sx = H->x_first ? ex : ex + 32, *sm = H->x_first ? ex + 32 : ex;
  const uint8_t *pkx = H->x_first ? pk : pk + eksz, *pkm = H->x_first ? pk + 32 : pk;
  uint8_t *ctx = H->x_first ? ct : ct + ctsz, *ctm = H->x_first ? ct + 32 : ct;
  uint8_t *ssx = H->x_first ? ss : ss + 32, *ssm = H->x_first ? ss + 32 : ss;
  int rc = 0;
  xkem_derive(ctx, esk, sx);
  if (!orc_x25519(ssx, esk, pkx)) rc = 1;
  if (H->mlkem) { if (orc_mlkem_encaps(H->k, ctm, ssm, pkm, sm)) rc = 1; }
  else orc_kyber_kem_encaps(H->k, ctm, ssm, pkm, sm);
  return rc;
}
/* Decapsulate (hybrid.go:263-283, xkem.go:185-200).  rc: 0 ok, 1 kem.ErrPubKey (low-order ct), 2 kem.ErrPrivKey */
int orc_hybrid_decaps(int id, uint8_t *ss, const uint8_t *sk, const uint8_t *ct) {
  const hyb *H = &HYB[id];
  const size_t dksz = orc_mlkem_dk_size(H->k), ctsz = orc_mlkem_ct_size(H->k);
  const uint8_t *skx = H->x_first ? sk : sk + dksz, *skm = H->x_first ? sk + 32 : sk;
  const uint8_t *ctx = H->x_first ? ct : ct + ctsz, *ctm = H->x_first ? ct + 32 : ct;
  uint8_t *ssx = H->x_first ? ss : ss + 32, *ssm = H->x_first ? ss + 32 : ss;
  int rc = 0;
  if (H->mlkem) { if (orc_mlkem_decaps(H->k, ssm, skm, ctm)) rc = 2; }
  else orc_kyber_kem_decaps(H->k, ssm, skm, ctm);
  if (!orc_x25519(ssx, skx, ctx)) rc = 1;
  return rc;
}
