// Host-side check of the "low format" fast path of csrc/kyber.cuh (compiled by nvcc as host code, no GPU needed):
//   1. the Shoup-form product mont_mul_lo equals the oracle's montReduce(zeta * b) for every twiddle (and the 1441 of
//      the inverse transform) and EVERY int16 b; barrett_lo equals barrettReduce on every int16;
//   2. an octet of eight lanes, emulated one lane after the other with the kernels' own pass / transposition functions,
//      reproduces orc_kyber_ntt / orc_kyber_invntt bit for bit on inputs inside the fast range, including the bounds;
//   3. words_in_range accepts exactly the inputs inside the bound.
// The kernels around these functions are covered by tests/test_gpu_kyber.py.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../circl_b200/csrc/kyber.cuh"

extern "C" {
void orc_kyber_ntt(int16_t p[256]);
void orc_kyber_invntt(int16_t p[256]);
int16_t orc_kyber_mont_reduce(int32_t x);
int16_t orc_kyber_barrett_reduce(int16_t x);
const int16_t* orc_kyber_zetas(void);
}
using namespace cb200::kyber;

static uint64_t st = 0x9e3779b97f4a7c15ull;
static uint64_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; }

// one polynomial through the fast path, lane by lane
static void fwd_fast(int16_t p[256], const TwLow* tab) {
  const uint32_t* words = reinterpret_cast<const uint32_t*>(p);
  alignas(16) unsigned char tile[kWideTileBytes];
  int32_t r[8][32];
  for (int v = 0; v < 8; v++) {
    for (int s = 0; s < 16; s++) unpack2_lo(words[8 * s + v], r[v][2 * s], r[v][2 * s + 1]);
    fwd_pass_S_lo(r[v]);
    wide_store_S(tile, v, r[v]);
  }
  uint32_t out[128];
  for (int v = 0; v < 8; v++) {
    wide_load_C(tile, v, r[v]);
    LaneTwLow t;
    load_lane_tw_lo(t, tab, v);
    fwd_pass_C_lo(r[v], t);
    for (int j = 0; j < 16; j++) out[16 * v + j] = pack2_lo(r[v][2 * j], r[v][2 * j + 1]);
  }
  memcpy(p, out, 512);
}
static void inv_fast(int16_t p[256], const TwLow* tab) {
  const uint32_t* words = reinterpret_cast<const uint32_t*>(p);
  alignas(16) unsigned char tile[kWideTileBytes];
  int32_t r[8][32];
  for (int v = 0; v < 8; v++) {
    for (int j = 0; j < 16; j++) unpack2_lo(words[16 * v + j], r[v][2 * j], r[v][2 * j + 1]);
    inv_pass_C_lo(r[v], tab, v);
    wide_store_C(tile, v, r[v]);
  }
  uint32_t out[128];
  for (int v = 0; v < 8; v++) {
    wide_load_S(tile, v, r[v]);
    inv_pass_S_lo(r[v], v);
    for (int s = 0; s < 16; s++) out[8 * s + v] = pack2_lo(r[v][2 * s], r[v][2 * s + 1]);
  }
  memcpy(p, out, 512);
}

int main() {
  int bad = 0;
  TwLow tab[128];
  for (int i = 0; i < 128; i++) tab[tw_slot(i)] = TwLow{zp_of(i), kk_of(i)};  // the device table order
  const int16_t* zt = orc_kyber_zetas();
  // 1. every twiddle x every int16
  for (int i = 0; i <= 128; i++) {
    const int32_t zeta = i < 128 ? zt[i] : 1441, zp = i < 128 ? tab[tw_slot(i)].zp : kScaleZp, kk = i < 128 ? tab[tw_slot(i)].kk : kScaleKk;
    if (i < 128 && zeta != zeta_of(i)) { bad++; printf("zeta table %d\n", i); }
    for (int b = -32768; b <= 32767; b++)
      if (mont_mul_lo(b, zp, kk) != orc_kyber_mont_reduce(zeta * b)) { bad++; if (bad < 10) printf("mont %d %d\n", i, b); }
  }
  for (int x = -32768; x <= 32767; x++)
    if (barrett_lo(x) != orc_kyber_barrett_reduce((int16_t)x)) { bad++; if (bad < 10) printf("barrett %d\n", x); }
  // 2. whole transforms inside the fast range
  for (int dir = 0; dir < 2; dir++) {
    const int bound = dir ? kInvBound : kFwdBound;
    for (int t = 0; t < 600; t++) {
      int16_t p[256], want[256];
      const int lim = t % 3 == 0 ? bound : (t % 3 == 1 ? 3329 : 1 + (int)(rnd() % bound));
      for (int i = 0; i < 256; i++) p[i] = (int16_t)((int)(rnd() % (2 * lim + 1)) - lim);
      if (t == 0) for (int i = 0; i < 256; i++) p[i] = (int16_t)bound;
      if (t == 3) for (int i = 0; i < 256; i++) p[i] = (int16_t)-bound;
      if (t == 6) for (int i = 0; i < 256; i++) p[i] = (int16_t)((i & 1) ? bound : -bound);
      if (t == 9) for (int i = 0; i < 256; i++) p[i] = (int16_t)((rnd() & 1) ? bound : -bound);
      memcpy(want, p, 512);
      uint32_t w16[16];
      bool in = true;
      for (int v = 0; v < 8; v++) {
        for (int s = 0; s < 16; s++) w16[s] = reinterpret_cast<const uint32_t*>(p)[8 * s + v];
        in = in && words_in_range(w16, bound);
      }
      if (!in) { bad++; printf("range predicate rejected an in-range polynomial (dir %d t %d)\n", dir, t); }
      if (dir) { orc_kyber_invntt(want); inv_fast(p, tab); } else { orc_kyber_ntt(want); fwd_fast(p, tab); }
      if (memcmp(p, want, 512)) { bad++; if (bad < 10) printf("transform mismatch dir %d t %d\n", dir, t); }
    }
  }
  // 3. the range predicate on single out-of-range coefficients
  for (int dir = 0; dir < 2; dir++) {
    const uint32_t bound = dir ? kInvBound : kFwdBound;
    for (int pos = 0; pos < 32; pos++)
      for (int val : {(int)bound + 1, -(int)bound - 1, 32767, -32768, (int)bound, -(int)bound}) {
        int16_t c[32] = {0};
        c[pos] = (int16_t)val;
        uint32_t w[16];
        memcpy(w, c, 64);
        const bool want = val >= -(int)bound && val <= (int)bound;
        if (words_in_range(w, bound) != want) { bad++; printf("range predicate dir %d pos %d val %d\n", dir, pos, val); }
      }
  }
  printf("bad=%d\n", bad);
  return bad != 0;
}
