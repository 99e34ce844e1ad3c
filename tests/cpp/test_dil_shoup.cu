// Host-side check (nvcc, no GPU) of the Shoup-form constant multiplication of csrc/dilithium.cuh: for every forward and
// inverse twiddle and ROver256, mont_mul_shoup(b, p, k) == montReduceLe2Q(c * b) of the oracle on edge values and a
// random sample of uint32 b (the identity is exact for EVERY b; see the comment at struct Shoup).
#include <cstdint>
#include <cstdio>

#include "../../circl_b200/csrc/dilithium.cuh"

extern "C" uint32_t orc_dil_mont_reduce_le2q(uint64_t x);
using namespace cb200::dil;

int main() {
  int bad = 0;
  uint64_t st = 0x2545F4914F6CDD1Dull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (uint32_t)(st >> 16); };
  const uint32_t edges[] = {0u, 1u, 2u, Q - 1, Q, Q + 1, 2 * Q - 1, 2 * Q, 256 * Q, 0x7fffffffu, 0x80000000u, 0x80000001u,
                            0xfffffffeu, 0xffffffffu};
  for (int i = 0; i <= 512; i++) {
    const uint32_t c = i < 256 ? zeta_of(i) : (i < 512 ? inv_zeta_of(i - 256) : ROVER256);
    const uint32_t p = shoup_p(c), k = shoup_k(c);
    if ((((uint64_t)p << 32) - c) != (uint64_t)k * Q || p >= Q) { bad++; printf("constants %d\n", i); }
    for (uint32_t b : edges)
      if (mont_mul_shoup(b, p, k) != orc_dil_mont_reduce_le2q((uint64_t)c * b)) { bad++; if (bad < 10) printf("edge %d %u\n", i, b); }
    for (int t = 0; t < 20000; t++) {
      const uint32_t b = rnd();
      if (mont_mul_shoup(b, p, k) != orc_dil_mont_reduce_le2q((uint64_t)c * b)) { bad++; if (bad < 10) printf("rnd %d %u\n", i, b); }
    }
  }
  // the lane-transposed shared-memory copy of the pair tables (stage_pairs): a permutation of 0..255 that puts "entry i
  // of lane v" of a layer (table position base + c v + i) at base + 8 i + v, the index the C-layout passes read
  for (int inv = 0; inv < 2; inv++) {
    bool seen[256] = {false};
    for (int p = 0; p < 256; p++) {
      const int d = inv ? staged_index<true>(p) : staged_index<false>(p);
      if (d < 0 || d > 255 || seen[d]) { bad++; printf("staged_index not a permutation (%d, %d)\n", inv, p); }
      else seen[d] = true;
    }
    const int bases[4] = {inv ? 0 : 128, inv ? 128 : 64, inv ? 192 : 32, inv ? 224 : 16}, per_lane[4] = {16, 8, 4, 2};
    for (int l = 0; l < 4; l++)
      for (int v = 0; v < 8; v++)
        for (int i = 0; i < per_lane[l]; i++) {
          const int table = bases[l] + per_lane[l] * v + i, want = bases[l] + 8 * i + v;
          const int got = inv ? staged_index<true>(table) : staged_index<false>(table);
          if (got != want) { bad++; printf("staged_index(%d) = %d, want %d\n", table, got, want); }
        }
  }
  printf("bad=%d\n", bad);
  return bad != 0;
}
