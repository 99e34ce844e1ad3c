/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the *generic* (purego) code path of
 * cloudflare/circl for the module-lattice hot path (ML-KEM / ML-DSA).
 * It is the parity checker for the CUDA product in circl_b200/ and the
 * "port" CPU baseline of bench.py.  Nothing under circl_b200/ may link,
 * import or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it.
 *
 * Pinning: see tests/test_oracle_*.py -- NIST ACVP vectors (FIPS 203/204),
 * PQCgenKAT transcript hashes, the reference's embedded sampler vectors and
 * Keccak KATs, all extracted from /root/reference by tests/golden/make_golden.py.
 *
 * All file:line citations are relative to the reference repository root.
 */
#ifndef CIRCL_B200_ORACLE_H
#define CIRCL_B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------- Keccak (internal/sha3) ---------------- */
void orc_keccak_f1600(uint64_t a[25]);                       /* keccakf.go:12 */
void orc_keccak_f1600_turbo(uint64_t a[25]);                 /* keccakf.go:12, turbo = true: rounds 12..23 */
typedef struct {
  uint64_t a[25];
  unsigned rate, pos;
  uint8_t ds;
  int squeezing;
} orc_sponge;
void orc_sponge_init(orc_sponge *s, unsigned rate, uint8_t ds); /* hashes.go:21,35; shake.go:56,74 */
void orc_sponge_write(orc_sponge *s, const uint8_t *p, size_t n); /* sha3.go:128 */
void orc_sponge_read(orc_sponge *s, uint8_t *out, size_t n);      /* sha3.go:163 */
void orc_sha3_256(uint8_t out[32], const uint8_t *in, size_t n);
void orc_sha3_512(uint8_t out[64], const uint8_t *in, size_t n);
void orc_shake128(uint8_t *out, size_t outlen, const uint8_t *in, size_t n);
void orc_shake256(uint8_t *out, size_t outlen, const uint8_t *in, size_t n);

/* ---------------- Kyber / ML-KEM (q = 3329) ---------------- */
#define ORC_KYBER_N 256
#define ORC_KYBER_Q 3329
int16_t orc_kyber_mont_reduce(int32_t x);     /* pke/kyber/internal/common/field.go:4-32 */
int16_t orc_kyber_barrett_reduce(int16_t x);  /* field.go:45-64 */
int16_t orc_kyber_csubq(int16_t x);           /* field.go:67-74 */
int16_t orc_kyber_to_mont(int16_t x);         /* field.go:35-39 */
const int16_t *orc_kyber_zetas(void);         /* ntt.go:16-29 (regenerated, not copied) */
void orc_kyber_ntt(int16_t p[256]);           /* ntt.go:60-135 */
void orc_kyber_invntt(int16_t p[256]);        /* ntt.go:145-193 */
void orc_kyber_mulhat(int16_t p[256], const int16_t a[256], const int16_t b[256]); /* poly.go:63-100 */
void orc_kyber_add(int16_t p[256], const int16_t a[256], const int16_t b[256]);
void orc_kyber_sub(int16_t p[256], const int16_t a[256], const int16_t b[256]);
void orc_kyber_barrett(int16_t p[256]);
void orc_kyber_normalize(int16_t p[256]);
void orc_kyber_tomont(int16_t p[256]);
void orc_kyber_pack(uint8_t buf[384], const int16_t p[256]);
void orc_kyber_unpack(int16_t p[256], const uint8_t buf[384]);
void orc_kyber_compress(uint8_t *m, const int16_t p[256], int d);
void orc_kyber_decompress(int16_t p[256], const uint8_t *m, int d);
void orc_kyber_msg_decompress(int16_t p[256], const uint8_t m[32]);
void orc_kyber_msg_compress(uint8_t m[32], const int16_t p[256]);
void orc_kyber_derive_noise(int16_t p[256], const uint8_t *seed, size_t seedlen, uint8_t nonce, int eta);
void orc_kyber_derive_uniform(int16_t p[256], const uint8_t seed[32], uint8_t x, uint8_t y);
/* batched helpers (n polynomials, contiguous) used by tests / cpu baseline */
void orc_kyber_ntt_batch(int16_t *p, size_t n, int inverse);
void orc_kyber_ntt_batch_mt(int16_t *p, size_t n, int inverse, int nthreads);
void orc_kyber_mulhat_batch(int16_t *p, const int16_t *a, const int16_t *b, size_t n);
void orc_kyber_dot_batch(int16_t *out, const int16_t *a, const int16_t *b, int k, size_t n);

/* ML-KEM (k = 2,3,4).  Sizes: ek = 384k+32, dk = 768k+96, ct = 32(du*k+dv) */
size_t orc_mlkem_ek_size(int k);
size_t orc_mlkem_dk_size(int k);
size_t orc_mlkem_ct_size(int k);
/* seed = d||z (64 B).  kem/mlkem/mlkem768/kyber.go:57-78 */
void orc_mlkem_keygen(int k, uint8_t *ek, uint8_t *dk, const uint8_t seed[64]);
/* returns 0, or -1 if ek is not canonical (kem.ErrPubKey, cpapke.go:45-55) */
int orc_mlkem_encaps(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *ek, const uint8_t m[32]);
/* returns 0, or -2 if H(ek) stored in dk mismatches (kem.ErrPrivKey) */
int orc_mlkem_decaps(int k, uint8_t ss[32], const uint8_t *dk, const uint8_t *ct);
/* round-3 Kyber512/768/1024 KEM (kem/kyber/kyber768/kyber.go:56-176): same K-PKE, different hashing */
void orc_kyber_kem_keygen(int k, uint8_t *ek, uint8_t *dk, const uint8_t seed[64]);
void orc_kyber_kem_encaps(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *ek, const uint8_t seed[32]);
void orc_kyber_kem_decaps(int k, uint8_t ss[32], const uint8_t *dk, const uint8_t *ct);
/* batched; ek_stride == 0 => one ek for all ops. returns number of failures */
/* ---- hybrid.c: X25519, X-Wing, kem/hybrid ---- */
int orc_x25519(uint8_t out[32], const uint8_t scalar[32], const uint8_t *point /* NULL = base point */);
void orc_xwing_keygen(uint8_t pk[1216], const uint8_t seed[32]);
int orc_xwing_encaps(uint8_t ct[1120], uint8_t ss[32], const uint8_t pk[1216], const uint8_t eseed[64]);
void orc_xwing_decaps(uint8_t ss[32], const uint8_t sk[32], const uint8_t ct[1120]);
size_t orc_hybrid_pk_size(int id); /* id 0 X25519MLKEM768, 1 Kyber768-X25519, 2 Kyber512-X25519 */
size_t orc_hybrid_sk_size(int id);
size_t orc_hybrid_ct_size(int id);
void orc_hybrid_keygen(int id, uint8_t *pk, uint8_t *sk, const uint8_t seed[64]);
int orc_hybrid_encaps(int id, uint8_t *ct, uint8_t *ss, const uint8_t *pk, const uint8_t seed[32]);
int orc_hybrid_decaps(int id, uint8_t *ss, const uint8_t *sk, const uint8_t *ct);

/* EncapsulateTo on an already unmarshalled key (its cached th, aT, hpk): BASELINE config 1, "pk pre-parsed" */
size_t orc_mlkem_parsed_size(void);
int orc_mlkem_pk_parse(int k, uint8_t *parsed, const uint8_t *ek);
void orc_mlkem_encaps_parsed(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *parsed, const uint8_t m[32]);
void orc_mlkem_encaps_parsed_batch(int k, uint8_t *ct, uint8_t *ss, const uint8_t *parsed, size_t stride, const uint8_t *m,
                                   size_t n, int nthreads);
int orc_mlkem_encaps_batch(int k, uint8_t *ct, uint8_t *ss, const uint8_t *ek, size_t ek_stride,
                           const uint8_t *m, size_t n, int nthreads);

/* ---- kyber_avx2.c: the reference's amd64 fast path restated with intrinsics (second CPU arm of bench.py) ----
 * f1600x4AVX2 (simd/keccakf1600/f1600x4_amd64.s:9), nttAVX2 / invNttAVX2 / mulHatAVX2 (pke/kyber/internal/common/amd64.s) */
void orc_keccak_f1600_x4(uint64_t *a /* 100 words: lane i of instance j at a[4 i + j] */);
void orc_kyber_ntt_avx2(int16_t p[256]);      /* == orc_kyber_ntt, coefficient for coefficient */
void orc_kyber_invntt_avx2(int16_t p[256]);   /* == orc_kyber_invntt modulo q (other lazy-reduction points) */
void orc_kyber_mulhat_avx2(int16_t p[256], const int16_t a[256], const int16_t b[256]); /* == orc_kyber_mulhat */
int orc_mlkem_encaps_avx2(int k, uint8_t *ct, uint8_t ss[32], const uint8_t *ek, const uint8_t m[32]);
int orc_mlkem_encaps_batch_avx2(int k, uint8_t *ct, uint8_t *ss, const uint8_t *ek, size_t ek_stride, const uint8_t *m,
                                size_t n, int nthreads);

/* ---------------- Dilithium / ML-DSA-65 (q = 8380417) ---------------- */
#define ORC_MLDSA65_PK 1952
#define ORC_MLDSA65_SK 4032
#define ORC_MLDSA65_SIG 3309
uint32_t orc_dil_mont_reduce_le2q(uint64_t x);        /* sign/internal/dilithium/field.go:20-24 */
const uint32_t *orc_dil_zetas(void);                   /* ntt.go:19 (regenerated) */
const uint32_t *orc_dil_inv_zetas(void);               /* ntt.go:66 (regenerated) */
void orc_dil_ntt(uint32_t p[256]);                     /* ntt.go:111-184 */
void orc_dil_invntt(uint32_t p[256]);                  /* ntt.go:191-217 */
void orc_dil_mulhat(uint32_t p[256], const uint32_t a[256], const uint32_t b[256]); /* poly.go:88 */
/* op: 0 add, 1 sub, 2 reduceLe2Q, 3 normalize, 4 normalizeAssumingLe2Q (poly.go:10-45) */
void orc_dil_poly_op(int op, uint32_t *p, const uint32_t *a, const uint32_t *b);
int orc_dil_exceeds(const uint32_t *p, uint32_t bound); /* poly.go:51-71 */
void orc_dil_decompose(const uint32_t *p, uint32_t *p0plusq, uint32_t *p1); /* mldsa65/internal/rounding.go:13-43 */
void orc_dil_ntt_batch(uint32_t *p, size_t n, int inverse);
void orc_dil_mulhat_batch(uint32_t *p, const uint32_t *a, const uint32_t *b, size_t n);
void orc_dil_derive_uniform(uint32_t p[256], const uint8_t seed[32], uint16_t nonce);   /* sample.go:92-123 */
void orc_dil_derive_leqeta(uint32_t p[256], const uint8_t seed[64], uint16_t nonce);    /* sample.go:129-181 */
void orc_dil_derive_legamma1(uint32_t p[256], const uint8_t seed[64], uint16_t nonce);  /* sample.go:197-209 */
void orc_dil_derive_ball(uint32_t p[256], const uint8_t seed[48]);                      /* sample.go:299-339 */
/* per-mode samplers (mode = 44, 65, 87) and the remaining leaf methods of generic.go */
void orc_mldsa_derive_leqeta(int mode, uint32_t p[256], const uint8_t seed[64], uint16_t nonce);
void orc_mldsa_derive_legamma1(int mode, uint32_t p[256], const uint8_t seed[64], uint16_t nonce);
void orc_mldsa_derive_ball(int mode, uint32_t p[256], const uint8_t *seed);     /* seed = c~: 32 / 48 / 64 bytes */
void orc_dil_power2round(const uint32_t *p, uint32_t *p0plusq, uint32_t *p1);  /* poly.go:77-84 */
void orc_dil_pack_le16(uint8_t buf[128], const uint32_t *p);                    /* pack.go:102-108 */
void orc_mldsa65_keygen(uint8_t pk[1952], uint8_t sk[4032], const uint8_t seed[32]);    /* internal/dilithium.go:181-241 */
/* returns the number of rejection-loop attempts (>= 1), or -1 after 576 (dilithium.go:372-377) */
int orc_mldsa65_sign(uint8_t sig[3309], const uint8_t *sk, const uint8_t *msg, size_t msglen, const uint8_t *ctx,
                     size_t ctxlen, const uint8_t rnd[32], int internal);
int orc_mldsa65_verify(const uint8_t pk[1952], const uint8_t *msg, size_t msglen, const uint8_t *ctx, size_t ctxlen,
                       const uint8_t *sig, size_t siglen, int internal);
int orc_mldsa65_sign_batch(uint8_t *sig, const uint8_t *sk, size_t sk_stride, const uint8_t *msgs, const uint64_t *off,
                           const uint8_t *rnd, size_t n, int nthreads);

/* run-time parameter set: mode = 44, 65 or 87 (sign/dilithium/gen.go:80-162) */
size_t orc_mldsa_sk_size(int mode);
size_t orc_mldsa_pk_size(int mode);
size_t orc_mldsa_sig_size(int mode);
void orc_mldsa_keygen(int mode, uint8_t *pk, uint8_t *sk, const uint8_t seed[32]);
int orc_mldsa_sign(int mode, uint8_t *sig, const uint8_t *sk, const uint8_t *msg, size_t msglen, const uint8_t *ctx,
                   size_t ctxlen, const uint8_t rnd[32], int internal);
int orc_mldsa_verify(int mode, const uint8_t *pk, const uint8_t *msg, size_t msglen, const uint8_t *ctx, size_t ctxlen,
                     const uint8_t *sig, size_t siglen, int internal);
int orc_mldsa_sign_batch(int mode, uint8_t *sig, const uint8_t *sk, size_t sk_stride, const uint8_t *msgs,
                         const uint64_t *off, const uint8_t *rnd, size_t n, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
