// Host-side check of the X25519 limb arithmetic of csrc/x25519.cuh against the oracle: the header is plain C++ once
// the CUDA qualifiers are defined away, so the field code is compared in the CPU suite (no GPU needed); the kernel
// around it is covered by tests/test_gpu_hybrid.py.
#include <cstdio>
#include <cstdint>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __noinline__
#include "../../circl_b200/csrc/x25519.cuh"
extern "C" int orc_x25519(uint8_t out[32], const uint8_t scalar[32], const uint8_t *point);
int main() {
  uint64_t st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
  int bad = 0;
  for (int t = 0; t < 300; t++) {
    uint32_t k[8], p[8], r[8];
    for (int i = 0; i < 8; i++) { k[i] = (uint32_t)rnd(); p[i] = (uint32_t)rnd(); }
    if (t % 7 == 0) { for (int i = 0; i < 8; i++) p[i] = 0xffffffffu; }          // non-canonical, bit 255 set
    if (t % 11 == 0) { memset(p, 0, 32); p[0] = t % 22 ? 1 : 0; }                // small order
    if (t == 5) { const uint8_t lo[32] = {0xec,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0xff,0x7f}; memcpy(p, lo, 32); }
    bool ok = cb200::x25519::scalarmult(r, k, p);
    uint8_t want[32];
    int wok = orc_x25519(want, (const uint8_t*)k, (const uint8_t*)p);
    if (memcmp(want, r, 32) || (int)ok != wok) { bad++; printf("mismatch t=%d ok=%d wok=%d\n", t, (int)ok, wok); }
  }
  // fixed-base path: table built on the host exactly as cb200_init does, results against the oracle's ladder
  static int32_t table[cb200::x25519::kBaseTableWords];
  cb200::x25519::build_base_table(table);
  for (int t = 0; t < 400; t++) {
    uint32_t k[8], r[8];
    for (int i = 0; i < 8; i++) k[i] = (uint32_t)rnd();
    if (t == 0) for (int i = 0; i < 8; i++) k[i] = 0xffffffffu;   // every digit carries; top nibble 7 + carry
    if (t == 1) for (int i = 0; i < 8; i++) k[i] = 0;             // clamp only
    if (t == 2) for (int i = 0; i < 8; i++) k[i] = 0x77777777u;
    if (t == 3) for (int i = 0; i < 8; i++) k[i] = 0x88888888u;
    cb200::x25519::scalarmult_base(r, k, table);
    uint8_t want[32];
    orc_x25519(want, (const uint8_t*)k, nullptr);
    if (memcmp(want, r, 32)) { bad++; printf("fixed-base mismatch t=%d\n", t); }
  }
  printf("bad=%d\n", bad);
  return bad != 0;
}
