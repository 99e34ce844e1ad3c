/*
 * oracle/kyber.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * CPU restatement of the reference's generic Kyber / ML-KEM path, standard
 * ("detangled") coefficient order:
 *   pke/kyber/internal/common/{field.go, ntt.go, poly.go, sample.go}
 *   pke/kyber/kyber768/internal/{vec.go, mat.go (non-X4 branch :14-29), cpapke.go}
 *   pke/kyber/kyber768/kyber.go:77-86 (ML-KEM keygen domain separation)
 *   kem/mlkem/mlkem768/kyber.go:57-78,103-137,144-184,247-263
 * K is a run-time parameter (2, 3, 4) instead of generated per-set packages.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>

#define N 256
#define Q 3329

/* ---- field.go ---- */
int16_t orc_kyber_mont_reduce(int32_t x) { /* field.go:4-32 */
  int16_t m = (int16_t)(uint32_t)((uint32_t)x * 62209u);
  return (int16_t)((uint32_t)(x - (int32_t)m * Q) >> 16);
}
int16_t orc_kyber_to_mont(int16_t x) { return orc_kyber_mont_reduce((int32_t)x * 1353); } /* field.go:35-39 */
int16_t orc_kyber_barrett_reduce(int16_t x) { /* field.go:45-64 */
  return (int16_t)(x - (int16_t)(((int32_t)x * 20159) >> 26) * Q);
}
int16_t orc_kyber_csubq(int16_t x) { /* field.go:67-74 */
  x = (int16_t)(x - Q);
  x = (int16_t)(x + ((x >> 15) & Q));
  return x;
}
#define mont orc_kyber_mont_reduce
#define barrett orc_kyber_barrett_reduce

/* ---- ntt.go: tables are regenerated from their defining formulas ---- */
static int16_t ZETAS[128];
static int zetas_ready;
static unsigned brv7(unsigned x) {
  unsigned r = 0;
  for (int i = 0; i < 7; i++) r |= ((x >> i) & 1u) << (6 - i);
  return r;
}
const int16_t *orc_kyber_zetas(void) { /* ntt.go:5-15: Zetas[i] = 17^brv7(i) * 2^16 mod q */
  if (!zetas_ready) {
    for (unsigned i = 0; i < 128; i++) {
      uint32_t z = 65536u % Q, e = brv7(i);
      for (uint32_t j = 0; j < e; j++) z = z * 17u % Q;
      ZETAS[i] = (int16_t)z;
    }
    zetas_ready = 1;
  }
  return ZETAS;
}

void orc_kyber_ntt(int16_t p[N]) { /* ntt.go:60-135 */
  const int16_t *zt = orc_kyber_zetas();
  int k = 0;
  for (int l = N / 2; l > 1; l >>= 1)
    for (int off = 0; off < N - l; off += 2 * l) {
      int32_t zeta = zt[++k];
      for (int j = off; j < off + l; j++) {
        int16_t t = mont(zeta * (int32_t)p[j + l]);
        p[j + l] = (int16_t)(p[j] - t);
        p[j] = (int16_t)(p[j] + t);
      }
    }
}

/*
 * Lazy Barrett schedule of the inverse NTT (ntt.go:38-50 InvNTTReductions).
 * Stated here as (layer, period, residues) instead of the flat index list:
 * after the layer with butterfly length l, coefficient i is reduced iff
 * (i mod period) is in the residue set.
 */
static int invntt_reduce_here(int l, int i) {
  switch (l) {
  case 8: { int r = i & 31; return r == 16 || r == 17; }                              /* after layer 3 */
  case 16: { int r = i & 63; return r == 0 || r == 1 || (r >= 32 && r <= 35); }       /* after layer 4 */
  case 32: { int r = i & 127; return r == 2 || r == 3 || (r >= 66 && r <= 71); }      /* after layer 5 */
  case 64: { return (i >= 4 && i <= 7) || (i >= 132 && i <= 143); }                   /* after layer 6 */
  default: return 0;
  }
}

void orc_kyber_invntt(int16_t p[N]) { /* ntt.go:145-193 */
  const int16_t *zt = orc_kyber_zetas();
  int k = 127;
  for (int l = 2; l < N; l <<= 1) {
    for (int off = 0; off < N - l; off += 2 * l) {
      int32_t minzeta = zt[k--];
      for (int j = off; j < off + l; j++) {
        int16_t t = (int16_t)(p[j + l] - p[j]);
        p[j] = (int16_t)(p[j] + p[j + l]);
        p[j + l] = mont(minzeta * (int32_t)t);
      }
    }
    for (int i = 0; i < N; i++)
      if (invntt_reduce_here(l, i)) p[i] = barrett(p[i]);
  }
  for (int j = 0; j < N; j++) p[j] = mont(1441 * (int32_t)p[j]);
}

/* ---- poly.go ---- */
void orc_kyber_mulhat(int16_t p[N], const int16_t a[N], const int16_t b[N]) { /* poly.go:63-100 */
  const int16_t *zt = orc_kyber_zetas();
  int k = 64;
  for (int i = 0; i < N; i += 4) {
    int32_t zeta = zt[k++];
    int16_t p0 = mont((int32_t)a[i + 1] * b[i + 1]);
    p0 = mont((int32_t)p0 * zeta);
    p0 = (int16_t)(p0 + mont((int32_t)a[i] * b[i]));
    int16_t p1 = mont((int32_t)a[i] * b[i + 1]);
    p1 = (int16_t)(p1 + mont((int32_t)a[i + 1] * b[i]));
    int16_t p2 = mont((int32_t)a[i + 3] * b[i + 3]);
    p2 = (int16_t)(-mont((int32_t)p2 * zeta));
    p2 = (int16_t)(p2 + mont((int32_t)a[i + 2] * b[i + 2]));
    int16_t p3 = mont((int32_t)a[i + 2] * b[i + 3]);
    p3 = (int16_t)(p3 + mont((int32_t)a[i + 3] * b[i + 2]));
    p[i] = p0; p[i + 1] = p1; p[i + 2] = p2; p[i + 3] = p3;
  }
}
void orc_kyber_add(int16_t p[N], const int16_t a[N], const int16_t b[N]) { for (int i = 0; i < N; i++) p[i] = (int16_t)(a[i] + b[i]); }
void orc_kyber_sub(int16_t p[N], const int16_t a[N], const int16_t b[N]) { for (int i = 0; i < N; i++) p[i] = (int16_t)(a[i] - b[i]); }
void orc_kyber_barrett(int16_t p[N]) { for (int i = 0; i < N; i++) p[i] = barrett(p[i]); }
void orc_kyber_normalize(int16_t p[N]) { for (int i = 0; i < N; i++) p[i] = orc_kyber_csubq(barrett(p[i])); }
void orc_kyber_tomont(int16_t p[N]) { for (int i = 0; i < N; i++) p[i] = orc_kyber_to_mont(p[i]); }

void orc_kyber_pack(uint8_t buf[384], const int16_t p[N]) { /* poly.go:106-116 */
  for (int i = 0; i < 128; i++) {
    uint16_t t0 = (uint16_t)p[2 * i], t1 = (uint16_t)p[2 * i + 1];
    buf[3 * i] = (uint8_t)t0;
    buf[3 * i + 1] = (uint8_t)((t0 >> 8) | (t1 << 4));
    buf[3 * i + 2] = (uint8_t)(t1 >> 4);
  }
}
void orc_kyber_unpack(int16_t p[N], const uint8_t buf[384]) { /* poly.go:123-129 */
  for (int i = 0; i < 128; i++) {
    p[2 * i] = (int16_t)(buf[3 * i] | ((buf[3 * i + 1] << 8) & 0xfff));
    p[2 * i + 1] = (int16_t)((buf[3 * i + 1] >> 4) | (buf[3 * i + 2] << 4));
  }
}

void orc_kyber_msg_decompress(int16_t p[N], const uint8_t m[32]) { /* poly.go:134-147 */
  for (int i = 0; i < 32; i++)
    for (int j = 0; j < 8; j++) p[8 * i + j] = (int16_t)(-(int16_t)((m[i] >> j) & 1) & ((Q + 1) / 2));
}
void orc_kyber_msg_compress(uint8_t m[32], const int16_t p[N]) { /* poly.go:150-166 */
  for (int i = 0; i < 32; i++) {
    m[i] = 0;
    for (int j = 0; j < 8; j++) {
      int16_t x = (int16_t)(1664 - p[8 * i + j]);
      x = (int16_t)((x >> 15) ^ x);
      x = (int16_t)(x - 832);
      m[i] |= (uint8_t)(((uint8_t)(x >> 15) & 1) << j);
    }
  }
}

/* Compress_q(x, d) per coefficient (poly.go:248-328), then a little-endian
 * d-bit stream -- identical bytes to the reference's hand-unrolled packers. */
static uint32_t compress1(int16_t x, int d) {
  uint32_t v = ((uint32_t)(int32_t)x << d) + Q / 2;
  if (d <= 5) return ((v * 315u) >> 20) & ((1u << d) - 1);
  return (uint32_t)(((uint64_t)v * 20642679ull) >> 36) & ((1u << d) - 1);
}
void orc_kyber_compress(uint8_t *m, const int16_t p[N], int d) {
  uint64_t acc = 0;
  int bits = 0, o = 0;
  for (int i = 0; i < N; i++) {
    acc |= (uint64_t)compress1(p[i], d) << bits;
    bits += d;
    while (bits >= 8) { m[o++] = (uint8_t)acc; acc >>= 8; bits -= 8; }
  }
}
void orc_kyber_decompress(int16_t p[N], const uint8_t *m, int d) { /* poly.go:170-243 */
  uint64_t acc = 0;
  int bits = 0, o = 0;
  for (int i = 0; i < N; i++) {
    while (bits < d) { acc |= (uint64_t)m[o++] << bits; bits += 8; }
    uint32_t t = (uint32_t)acc & ((1u << d) - 1);
    acc >>= d; bits -= d;
    p[i] = (int16_t)(((1u << (d - 1)) + t * Q) >> d);
  }
}

/* ---- sample.go ---- */
void orc_kyber_derive_noise(int16_t p[N], const uint8_t *seed, size_t seedlen, uint8_t nonce, int eta) {
  orc_sponge h;
  uint8_t buf[192 + 8] = {0};
  orc_sponge_init(&h, 136, 0x1f);
  orc_sponge_write(&h, seed, seedlen);
  orc_sponge_write(&h, &nonce, 1);
  if (eta == 2) { /* sample.go:67-95 */
    orc_sponge_read(&h, buf, 128);
    for (int i = 0; i < 16; i++) {
      uint64_t t; memcpy(&t, buf + 8 * i, 8);
      uint64_t d = t & 0x5555555555555555ull;
      d += (t >> 1) & 0x5555555555555555ull;
      for (int j = 0; j < 16; j++) {
        int16_t a = (int16_t)(d & 3); d >>= 2;
        int16_t b = (int16_t)(d & 3); d >>= 2;
        p[16 * i + j] = (int16_t)(a - b);
      }
    }
  } else { /* eta == 3, sample.go:31-62 */
    orc_sponge_read(&h, buf, 192);
    for (int i = 0; i < 32; i++) {
      uint64_t t = 0; memcpy(&t, buf + 6 * i, 8);
      uint64_t d = t & 0x249249249249ull;
      d += (t >> 1) & 0x249249249249ull;
      d += (t >> 2) & 0x249249249249ull;
      for (int j = 0; j < 8; j++) {
        int16_t a = (int16_t)(d & 7); d >>= 3;
        int16_t b = (int16_t)(d & 7); d >>= 3;
        p[8 * i + j] = (int16_t)(a - b);
      }
    }
  }
}

void orc_kyber_derive_uniform(int16_t p[N], const uint8_t seed[32], uint8_t x, uint8_t y) { /* sample.go:192-236 */
  orc_sponge h;
  uint8_t buf[168], suffix[2] = {x, y};
  orc_sponge_init(&h, 168, 0x1f);
  orc_sponge_write(&h, seed, 32);
  orc_sponge_write(&h, suffix, 2);
  int i = 0;
  while (i < N) {
    orc_sponge_read(&h, buf, 168);
    for (int j = 0; j < 168 && i < N; j += 3) {
      uint16_t t1 = (uint16_t)((buf[j] | (buf[j + 1] << 8)) & 0xfff);
      uint16_t t2 = (uint16_t)(((buf[j + 1] >> 4) | (buf[j + 2] << 4)) & 0xfff);
      if (t1 < Q) p[i++] = (int16_t)t1;
      if (i < N && t2 < Q) p[i++] = (int16_t)t2;
    }
  }
}

/* ---- batched helpers ---- */
void orc_kyber_ntt_batch(int16_t *p, size_t n, int inverse) {
  for (size_t i = 0; i < n; i++) inverse ? orc_kyber_invntt(p + i * N) : orc_kyber_ntt(p + i * N);
}
void orc_kyber_mulhat_batch(int16_t *p, const int16_t *a, const int16_t *b, size_t n) {
  for (size_t i = 0; i < n; i++) orc_kyber_mulhat(p + i * N, a + i * N, b + i * N);
}
/* PolyDotHat (vec.go:30-37): out = sum_{i<k} MulHat(a[i], b[i]), unreduced int16 adds */
static void dot_hat(int16_t out[N], const int16_t *a, const int16_t *b, int k) {
  int16_t t[N];
  memset(out, 0, N * sizeof(int16_t));
  for (int i = 0; i < k; i++) {
    orc_kyber_mulhat(t, a + i * N, b + i * N);
    orc_kyber_add(out, t, out);
  }
}
void orc_kyber_dot_batch(int16_t *out, const int16_t *a, const int16_t *b, int k, size_t n) {
  for (size_t i = 0; i < n; i++) dot_hat(out + i * N, a + i * k * N, b + i * k * N, k);
}

/* ---- parameter sets (pke/kyber/kyber{512,768,1024}/internal/params.go) ---- */
static int eta1_of(int k) { return k == 2 ? 3 : 2; }
static int du_of(int k) { return k == 4 ? 11 : 10; }
static int dv_of(int k) { return k == 4 ? 5 : 4; }
size_t orc_mlkem_ek_size(int k) { return 384u * k + 32; }
size_t orc_mlkem_dk_size(int k) { return 768u * k + 96; }
size_t orc_mlkem_ct_size(int k) { return 32u * (du_of(k) * k + dv_of(k)); }

/* Mat.Derive (mat.go:13-29). aT[i][j] = XOF(rho, i, j) if transpose else XOF(rho, j, i) */
static void mat_derive(int16_t *m, int k, const uint8_t rho[32], int transpose) {
  for (int i = 0; i < k; i++)
    for (int j = 0; j < k; j++)
      orc_kyber_derive_uniform(m + (i * k + j) * N, rho, (uint8_t)(transpose ? i : j), (uint8_t)(transpose ? j : i));
}

/* K-PKE.Encrypt, cpapke.go:137-181. th is normalized NTT(t), aT the transposed matrix */
static void cpapke_encrypt(int k, uint8_t *ct, const int16_t *th, const int16_t *aT, const uint8_t pt[32],
                           const uint8_t seed[32]) {
  int16_t rh[4 * N], e1[4 * N], u[4 * N], e2[N], v[N], m[N];
  int du = du_of(k), dv = dv_of(k);
  for (int i = 0; i < k; i++) orc_kyber_derive_noise(rh + i * N, seed, 32, (uint8_t)i, eta1_of(k));
  for (int i = 0; i < k; i++) { orc_kyber_ntt(rh + i * N); orc_kyber_barrett(rh + i * N); }
  for (int i = 0; i < k; i++) orc_kyber_derive_noise(e1 + i * N, seed, 32, (uint8_t)(k + i), 2);
  orc_kyber_derive_noise(e2, seed, 32, (uint8_t)(2 * k), 2);
  for (int i = 0; i < k; i++) dot_hat(u + i * N, aT + i * k * N, rh, k);
  for (int i = 0; i < k; i++) {
    orc_kyber_barrett(u + i * N);
    orc_kyber_invntt(u + i * N);
    orc_kyber_add(u + i * N, u + i * N, e1 + i * N);
  }
  dot_hat(v, th, rh, k);
  orc_kyber_barrett(v);
  orc_kyber_invntt(v);
  orc_kyber_msg_decompress(m, pt);
  orc_kyber_add(v, v, m);
  orc_kyber_add(v, v, e2);
  for (int i = 0; i < k; i++) {
    orc_kyber_normalize(u + i * N);
    orc_kyber_compress(ct + i * 32 * du, u + i * N, du);
  }
  orc_kyber_normalize(v);
  orc_kyber_compress(ct + k * 32 * du, v, dv);
}

/* PublicKey.UnpackMLKEM (cpapke.go:45-63): unpack, normalize, modulus check, derive aT */
static int pk_unpack(int k, int16_t *th, int16_t *aT, const uint8_t *ek) {
  uint8_t chk[384];
  int bad = 0;
  for (int i = 0; i < k; i++) {
    orc_kyber_unpack(th + i * N, ek + 384 * i);
    orc_kyber_normalize(th + i * N);
    orc_kyber_pack(chk, th + i * N);
    if (memcmp(chk, ek + 384 * i, 384)) bad = 1;
  }
  mat_derive(aT, k, ek + 384 * k, 1);
  return bad ? -1 : 0;
}

static void kem_keygen(int k, int mlkem, uint8_t *ek, uint8_t *dk, const uint8_t seed[64]) {
  /* kem/mlkem/mlkem768/kyber.go:57-78 -> pke/kyber/kyber768/kyber.go:77-86 -> cpapke.go:66-110;
   * round-3 Kyber (kem/kyber/kyber768/kyber.go:56-77) hashes the 32-byte seed without the K byte */
  uint8_t seed2[33], exp[64];
  int16_t A[16 * N], sh[4 * N], eh[4 * N], th[4 * N];
  memcpy(seed2, seed, 32);
  seed2[32] = (uint8_t)k;
  orc_sha3_512(exp, seed2, mlkem ? 33 : 32);
  const uint8_t *rho = exp, *sigma = exp + 32;
  mat_derive(A, k, rho, 0);
  for (int i = 0; i < k; i++) {
    orc_kyber_derive_noise(sh + i * N, sigma, 32, (uint8_t)i, eta1_of(k));
    orc_kyber_ntt(sh + i * N);
    orc_kyber_normalize(sh + i * N);
  }
  for (int i = 0; i < k; i++) {
    orc_kyber_derive_noise(eh + i * N, sigma, 32, (uint8_t)(k + i), eta1_of(k));
    orc_kyber_ntt(eh + i * N);
  }
  for (int i = 0; i < k; i++) {
    dot_hat(th + i * N, A + i * k * N, sh, k);
    orc_kyber_tomont(th + i * N);
    orc_kyber_add(th + i * N, th + i * N, eh + i * N);
    orc_kyber_normalize(th + i * N);
  }
  for (int i = 0; i < k; i++) orc_kyber_pack(ek + 384 * i, th + i * N);
  memcpy(ek + 384 * k, rho, 32);
  /* dk = sk || ek || H(ek) || z  (kyber.go:187-201) */
  for (int i = 0; i < k; i++) orc_kyber_pack(dk + 384 * i, sh + i * N);
  size_t eksz = orc_mlkem_ek_size(k);
  memcpy(dk + 384 * k, ek, eksz);
  orc_sha3_256(dk + 384 * k + eksz, ek, eksz);
  memcpy(dk + 384 * k + eksz + 32, seed + 32, 32);
}

void orc_mlkem_keygen(int k, uint8_t *ek, uint8_t *dk, const uint8_t seed[64]) { kem_keygen(k, 1, ek, dk, seed); }
void orc_kyber_kem_keygen(int k, uint8_t *ek, uint8_t *dk, const uint8_t seed[64]) { kem_keygen(k, 0, ek, dk, seed); }

/* PublicKey.Unpack (cpapke.go:58-63): no modulus check, t-hat normalised */
static void pk_unpack_lenient(int k, int16_t *th, int16_t *aT, const uint8_t *ek) {
  for (int i = 0; i < k; i++) {
    orc_kyber_unpack(th + i * N, ek + 384 * i);
    orc_kyber_normalize(th + i * N);
  }
  mat_derive(aT, k, ek + 384 * k, 1);
}

/* round-3 Kyber Encapsulate (kem/kyber/kyber768/kyber.go:98-137): m = H(seed), (K, r) = G(m || H(pk)),
 * ct = Enc(pk, m, r), ss = KDF(K || H(ct)) */
void orc_kyber_kem_encaps(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *ek, const uint8_t seed[32]) {
  int16_t th[4 * N], aT[16 * N];
  uint8_t m[32], g_in[64], kr[64];
  size_t ctsz = orc_mlkem_ct_size(k);
  pk_unpack_lenient(k, th, aT, ek);
  orc_sha3_256(m, seed, 32);
  memcpy(g_in, m, 32);
  orc_sha3_256(g_in + 32, ek, orc_mlkem_ek_size(k));
  orc_sha3_512(kr, g_in, 64);
  cpapke_encrypt(k, ct, th, aT, m, kr + 32);
  orc_sha3_256(kr + 32, ct, ctsz);
  orc_shake256(ss, 32, kr, 64);
}

/* round-3 Kyber Decapsulate (kem/kyber/kyber768/kyber.go:139-176) */
void orc_kyber_kem_decaps(int k, uint8_t ss[32], const uint8_t *dk, const uint8_t *ct) {
  int16_t sh[4 * N], th[4 * N], aT[16 * N], u[4 * N], v[N], mp[N];
  size_t eksz = orc_mlkem_ek_size(k), ctsz = orc_mlkem_ct_size(k);
  const uint8_t *ek = dk + 384 * k, *hpk = ek + eksz, *z = hpk + 32;
  uint8_t m2[32], g_in[64], kr2[64];
  int du = du_of(k), dv = dv_of(k);
  for (int i = 0; i < k; i++) { orc_kyber_unpack(sh + i * N, dk + 384 * i); orc_kyber_normalize(sh + i * N); }
  pk_unpack_lenient(k, th, aT, ek);
  for (int i = 0; i < k; i++) { orc_kyber_decompress(u + i * N, ct + i * 32 * du, du); orc_kyber_ntt(u + i * N); }
  orc_kyber_decompress(v, ct + k * 32 * du, dv);
  dot_hat(mp, sh, u, k);
  orc_kyber_barrett(mp);
  orc_kyber_invntt(mp);
  orc_kyber_sub(mp, v, mp);
  orc_kyber_normalize(mp);
  orc_kyber_msg_compress(m2, mp);
  memcpy(g_in, m2, 32);
  memcpy(g_in + 32, hpk, 32);
  orc_sha3_512(kr2, g_in, 64);
  uint8_t *ct2 = (uint8_t *)malloc(ctsz);
  cpapke_encrypt(k, ct2, th, aT, m2, kr2 + 32);
  orc_sha3_256(kr2 + 32, ct, ctsz);
  if (memcmp(ct, ct2, ctsz) != 0) memcpy(kr2, z, 32);
  free(ct2);
  orc_shake256(ss, 32, kr2, 64);
}

int orc_mlkem_encaps(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *ek, const uint8_t m[32]) {
  /* scheme.UnmarshalBinaryPublicKey (kyber.go:247-263) then EncapsulateTo (kyber.go:103-137) */
  int16_t th[4 * N], aT[16 * N];
  uint8_t g_in[64], kr[64];
  int rc = pk_unpack(k, th, aT, ek);
  if (rc) return rc;
  memcpy(g_in, m, 32);
  orc_sha3_256(g_in + 32, ek, orc_mlkem_ek_size(k));
  orc_sha3_512(kr, g_in, 64);
  cpapke_encrypt(k, ct, th, aT, m, kr + 32);
  memcpy(ss, kr, 32);
  return 0;
}

int orc_mlkem_decaps(int k, uint8_t ss[32], const uint8_t *dk, const uint8_t *ct) {
  /* PrivateKey.Unpack (kyber.go:203-229) + DecapsulateTo (kyber.go:144-184) + DecryptTo (cpapke.go:113-130) */
  int16_t sh[4 * N], th[4 * N], aT[16 * N], u[4 * N], v[N], mp[N];
  size_t eksz = orc_mlkem_ek_size(k), ctsz = orc_mlkem_ct_size(k);
  const uint8_t *ek = dk + 384 * k, *hpk = ek + eksz, *z = hpk + 32;
  uint8_t h2[32], m2[32], g_in[64], kr2[64], ss2[32];
  int du = du_of(k), dv = dv_of(k);
  for (int i = 0; i < k; i++) { orc_kyber_unpack(sh + i * N, dk + 384 * i); orc_kyber_normalize(sh + i * N); }
  /* note: dk-embedded ek goes through Unpack (no modulus check), cpapke.go:58-63 */
  for (int i = 0; i < k; i++) { orc_kyber_unpack(th + i * N, ek + 384 * i); orc_kyber_normalize(th + i * N); }
  mat_derive(aT, k, ek + 384 * k, 1);
  orc_sha3_256(h2, ek, eksz);
  if (memcmp(h2, hpk, 32)) return -2;
  for (int i = 0; i < k; i++) { orc_kyber_decompress(u + i * N, ct + i * 32 * du, du); orc_kyber_ntt(u + i * N); }
  orc_kyber_decompress(v, ct + k * 32 * du, dv);
  dot_hat(mp, sh, u, k);
  orc_kyber_barrett(mp);
  orc_kyber_invntt(mp);
  orc_kyber_sub(mp, v, mp);
  orc_kyber_normalize(mp);
  orc_kyber_msg_compress(m2, mp);
  memcpy(g_in, m2, 32);
  memcpy(g_in + 32, hpk, 32);
  orc_sha3_512(kr2, g_in, 64);
  uint8_t *ct2 = (uint8_t *)malloc(ctsz);
  cpapke_encrypt(k, ct2, th, aT, m2, kr2 + 32);
  orc_sponge prf;
  orc_sponge_init(&prf, 136, 0x1f);
  orc_sponge_write(&prf, z, 32);
  orc_sponge_write(&prf, ct, ctsz);
  orc_sponge_read(&prf, ss2, 32);
  int same = memcmp(ct, ct2, ctsz) == 0;
  free(ct2);
  memcpy(ss, same ? kr2 : ss2, 32);
  return 0;
}

#include <pthread.h>
typedef struct {
  int k; uint8_t *ct, *ss; const uint8_t *ek; size_t ek_stride; const uint8_t *m; size_t lo, hi; int fails;
} enc_job;
static void *enc_worker(void *arg) {
  enc_job *j = (enc_job *)arg;
  size_t ctsz = orc_mlkem_ct_size(j->k);
  for (size_t i = j->lo; i < j->hi; i++)
    if (orc_mlkem_encaps(j->k, j->ct + i * ctsz, j->ss + i * 32, j->ek + i * j->ek_stride, j->m + i * 32)) j->fails++;
  return NULL;
}
int orc_mlkem_encaps_batch(int k, uint8_t *ct, uint8_t *ss, const uint8_t *ek, size_t ek_stride, const uint8_t *m,
                           size_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  enc_job *jobs = (enc_job *)malloc(sizeof(enc_job) * nthreads);
  int fails = 0;
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (enc_job){k, ct, ss, ek, ek_stride, m, n * t / nthreads, n * (t + 1) / nthreads, 0};
    pthread_create(&th[t], NULL, enc_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); fails += jobs[t].fails; }
  free(th); free(jobs);
  return fails;
}

/* ---- "pk pre-parsed" variant for BASELINE config 1: the reference's BenchmarkEncapsulate (kem/schemes/schemes_test.go)
 * and the KAT loop (kem/kyber/kat_test.go:48-81) time EncapsulateTo on an already unmarshalled *PublicKey, whose cached
 * fields are th, aT and hpk (kem/mlkem/mlkem768/kyber.go:39-43, cpapke.go:19-25).  parsed = th[4][256] | aT[16][256]
 * (int16) | hpk[32]. */
size_t orc_mlkem_parsed_size(void) { return 20 * N * sizeof(int16_t) + 32; }
int orc_mlkem_pk_parse(int k, uint8_t *parsed, const uint8_t *ek) {
  int16_t *th = (int16_t *)parsed, *aT = th + 4 * N;
  int rc = pk_unpack(k, th, aT, ek);
  orc_sha3_256(parsed + 20 * N * sizeof(int16_t), ek, orc_mlkem_ek_size(k));
  return rc;
}
void orc_mlkem_encaps_parsed(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *parsed, const uint8_t m[32]) {
  const int16_t *th = (const int16_t *)parsed, *aT = th + 4 * N;
  uint8_t g_in[64], kr[64];
  memcpy(g_in, m, 32);
  memcpy(g_in + 32, parsed + 20 * N * sizeof(int16_t), 32);
  orc_sha3_512(kr, g_in, 64);
  cpapke_encrypt(k, ct, th, aT, m, kr + 32);
  memcpy(ss, kr, 32);
}
typedef struct {
  int k; uint8_t *ct, *ss; const uint8_t *parsed; size_t stride; const uint8_t *m; size_t lo, hi;
} encp_job;
static void *encp_worker(void *arg) {
  encp_job *j = (encp_job *)arg;
  size_t ctsz = orc_mlkem_ct_size(j->k);
  for (size_t i = j->lo; i < j->hi; i++)
    orc_mlkem_encaps_parsed(j->k, j->ct + i * ctsz, j->ss + i * 32, j->parsed + i * j->stride, j->m + i * 32);
  return NULL;
}
void orc_mlkem_encaps_parsed_batch(int k, uint8_t *ct, uint8_t *ss, const uint8_t *parsed, size_t stride, const uint8_t *m,
                                   size_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  encp_job *jobs = (encp_job *)malloc(sizeof(encp_job) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (encp_job){k, ct, ss, parsed, stride, m, n * t / nthreads, n * (t + 1) / nthreads};
    pthread_create(&th[t], NULL, encp_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

typedef struct { int16_t *p; size_t lo, hi; int inverse; } ntt_job;
static void *ntt_worker(void *arg) {
  ntt_job *j = (ntt_job *)arg;
  orc_kyber_ntt_batch(j->p + j->lo * N, j->hi - j->lo, j->inverse);
  return NULL;
}
/* multi-threaded batch NTT for the CPU baseline leg of bench.py */
void orc_kyber_ntt_batch_mt(int16_t *p, size_t n, int inverse, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > n) nthreads = n ? (int)n : 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  ntt_job *jobs = (ntt_job *)malloc(sizeof(ntt_job) * nthreads);
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = (ntt_job){p, n * t / nthreads, n * (t + 1) / nthreads, inverse};
    pthread_create(&th[t], NULL, ntt_worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}
