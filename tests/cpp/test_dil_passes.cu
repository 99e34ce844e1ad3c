// Host-side check (nvcc, no GPU) of the Dilithium transform passes of csrc/dilithium.cuh: an octet of eight lanes is
// emulated one lane after the other with the kernels' own pass functions (S-layout pass with immediate Shoup pairs,
// transposition through the padded tile, C-layout pass with the pairs read from the lane-transposed staged copy) and
// compared with the oracle's nttGeneric / invNttGeneric restatement (sign/internal/dilithium/ntt.go:111-217),
// unnormalised, on arbitrary uint32 inputs.  The kernels around these functions are covered by tests/test_gpu_dilithium.py.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../circl_b200/csrc/dilithium.cuh"

extern "C" void orc_dil_ntt_batch(uint32_t* p, size_t n, int inverse);
using namespace cb200::dil;

static uint64_t st = 0x9e3779b97f4a7c15ull;
static uint32_t rnd() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 13); }

static void stage(uint2* dst, bool inv) {  // what stage_pairs<INV> writes
  for (int q = 0; q < 256; q++) {
    const uint32_t c = inv ? inv_zeta_of(q) : zeta_of(q);
    uint2 z;
    z.x = shoup_p(c);
    z.y = shoup_k(c);
    dst[inv ? staged_index<true>(q) : staged_index<false>(q)] = z;
  }
}
static void fwd(uint32_t p[256], const uint2* zs) {
  alignas(16) uint32_t tile[kPolyWords];
  uint32_t r[8][32];
  for (int v = 0; v < 8; v++) {
    for (int s = 0; s < 16; s++) { r[v][2 * s] = p[16 * s + 2 * v]; r[v][2 * s + 1] = p[16 * s + 2 * v + 1]; }
    fwd_pass_S(r[v]);
    store_S(tile, v, r[v]);
  }
  for (int v = 0; v < 8; v++) {
    load_C(tile, v, r[v]);
    fwd_pass_C_smem(r[v], zs, v);
    memcpy(p + 32 * v, r[v], 128);
  }
}
static void inv(uint32_t p[256], const uint2* iz) {
  alignas(16) uint32_t tile[kPolyWords];
  uint32_t r[8][32];
  for (int v = 0; v < 8; v++) {
    memcpy(r[v], p + 32 * v, 128);
    inv_pass_C_smem(r[v], iz, v);
    store_C(tile, v, r[v]);
  }
  for (int v = 0; v < 8; v++) {
    load_S(tile, v, r[v]);
    inv_pass_S(r[v]);
    for (int s = 0; s < 16; s++) { p[16 * s + 2 * v] = r[v][2 * s]; p[16 * s + 2 * v + 1] = r[v][2 * s + 1]; }
  }
}

int main() {
  int bad = 0;
  static uint2 zs[256], iz[256];
  stage(zs, false);
  stage(iz, true);
  for (int dir = 0; dir < 2; dir++)
    for (int t = 0; t < 300; t++) {
      uint32_t p[256], want[256];
      for (int i = 0; i < 256; i++) p[i] = t % 3 == 0 ? rnd() % Q : (t % 3 == 1 ? rnd() % (2 * Q) : rnd());
      if (t == 0) for (int i = 0; i < 256; i++) p[i] = Q - 1;
      if (t == 3) for (int i = 0; i < 256; i++) p[i] = 0xffffffffu;
      memcpy(want, p, sizeof p);
      orc_dil_ntt_batch(want, 1, dir);
      if (dir) inv(p, iz); else fwd(p, zs);
      if (memcmp(p, want, sizeof p)) { bad++; if (bad < 10) printf("transform mismatch dir %d t %d\n", dir, t); }
    }
  printf("bad=%d\n", bad);
  return bad != 0;
}
