// circl_b200/csrc/kyber_kernels.cu -- batched q=3329 kernels (sm_100a) and their launchers.
//
// cb200_kyber_ntt      <- (*Poly).NTT / InvNTT      pke/kyber/internal/common/generic.go:24,36 (stubs_amd64.go:8-14)
// cb200_kyber_mulhat   <- (*Poly).MulHat            generic.go:49 (stubs_amd64.go:17)
// cb200_kyber_dot      <- PolyDotHat                pke/kyber/kyber768/internal/vec.go:30-37
// cb200_kyber_poly_op  <- Add/Sub/BarrettReduce/Normalize/ToMont  generic.go:7-77, poly.go:48
#include <stdlib.h>

#include "context.h"
#include "kyber.cuh"
#include "launch.h"

namespace cb200 {
namespace kyber {

constexpr int kThreads = 128;               // 4 warps = 16 octets = 16 polynomials in flight per CTA
constexpr int kOctetsPerCta = kThreads / 8;

// In-place forward / inverse NTT over a batch.  HBM traffic: 512 B read + 512 B
// written per polynomial (the algorithmic minimum).  The 1 KiB twiddle table is staged
// into shared memory once per CTA by a bulk-async (TMA, UBLKCP) copy.
template <bool INV>
__global__ void __launch_bounds__(kThreads, INV ? 8 : 6) ntt_kernel(uint32_t* __restrict__ polys, size_t n,
                                                          const TwPair* __restrict__ tw) {
  __shared__ __align__(16) uint32_t tiles[kOctetsPerCta * kPolyWords];
  __shared__ __align__(16) TwPair tws[128];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int oct = lane >> 3, v = lane & 7;
  uint32_t* tile = tiles + (warp * 4 + oct) * kPolyWords;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, sizeof(tws));
    bulk_g2s(tws, tw, sizeof(tws), &bar);
  }
  mbar_wait(&bar, 0);
  const volatile TwPair* tab = tws;
  // Measured on B200 (profiles/): the forward transform is fastest with its 14 per-lane twiddle pairs held in
  // registers (80 regs, 6 CTAs/SM), the inverse with the twiddles read from shared memory at the point of
  // use (64 regs, 8 CTAs/SM).
  LaneTw t;
  if (!INV) load_lane_tw(t, tws, v);

  const size_t stride = (size_t)gridDim.x * kOctetsPerCta;
  for (size_t base = ((size_t)blockIdx.x * 4 + warp) * 4; base < n; base += stride) {
    const size_t p = base + oct;
    const bool active = p < n;
    uint32_t* poly = polys + (active ? p : n - 1) * (N / 2);  // idle octets recompute the last polynomial, never store
    int32_t r[32];
    if (!INV) {
      gload_S(poly, v, r);
      fwd_pass_S(r);
      store_S(tile, v, r);
      __syncwarp();
      load_C(tile, v, r);
      fwd_pass_C(r, t);
      if (active) gstore_C(poly, v, r);
    } else {
      gload_C(poly, v, r);
      inv_pass_C_smem(r, tab, v);
      store_C(tile, v, r);
      __syncwarp();
      load_S(tile, v, r);
      inv_pass_S(r, v);
      if (active) gstore_S(poly, v, r);
    }
    __syncwarp();
  }
}

// out[i] = sum_{j<k} MulHat(a[i*k+j], b[i*k+j])   (k = 1: plain MulHat)
__global__ void __launch_bounds__(kThreads) dot_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ a,
                                                       const uint32_t* __restrict__ b, int k, size_t n,
                                                       const TwPair* __restrict__ tw) {
  const int v = threadIdx.x & 7;
  LaneTw t;
  load_lane_tw(t, tw, v);
  const size_t stride = (size_t)gridDim.x * kOctetsPerCta;
  for (size_t p = (size_t)blockIdx.x * kOctetsPerCta + (threadIdx.x >> 3); p < n; p += stride) {
    int32_t acc[32];
#pragma unroll
    for (int i = 0; i < 32; i++) acc[i] = 0;
    for (int j = 0; j < k; j++) {
      int32_t x[32], y[32];
      gload_C(a + (p * k + j) * (N / 2), v, x);
      gload_C(b + (p * k + j) * (N / 2), v, y);
      mulhat_acc_C(acc, x, y, t);
    }
    gstore_C(out + p * (N / 2), v, acc);
  }
}

// element-wise family; one thread handles 8 coefficients (one 128-bit access)
enum PolyOp { OP_ADD = 0, OP_SUB = 1, OP_BARRETT = 2, OP_NORMALIZE = 3, OP_TOMONT = 4 };
template <int OP>
__device__ __forceinline__ int32_t ew(int32_t x, int32_t y) {
  if (OP == OP_ADD) return x + y;
  if (OP == OP_SUB) return x - y;
  if (OP == OP_BARRETT) return barrett_hi(x);
  if (OP == OP_NORMALIZE) return csubq_hi(barrett_hi(x));
  return mont_mul_hi(x >> 16, 1353, (int32_t)(((1353u * QINV) & 0xffffu) << 16));  // toMont, field.go:35-39
}
template <int OP>
__global__ void __launch_bounds__(256) ew_kernel(uint4* __restrict__ out, const uint4* __restrict__ a,
                                                 const uint4* __restrict__ b, size_t nvec) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    uint4 x = ldg_stream128(a + i), y = make_uint4(0, 0, 0, 0);
    if (OP == OP_ADD || OP == OP_SUB) y = ldg_stream128(b + i);
    uint32_t xs[4] = {x.x, x.y, x.z, x.w}, ys[4] = {y.x, y.y, y.z, y.w}, o[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
      int32_t xl, xh, yl, yh;
      unpack2(xs[w], xl, xh);
      unpack2(ys[w], yl, yh);
      o[w] = pack2(ew<OP>(xl, yl), ew<OP>(xh, yh));
    }
    stg_stream128(out + i, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// Forward NTT with the input staged by the TMA engine.  Every octet owns two 608-byte slots of shared memory; lane 0
// of the octet bulk-copies the next polynomial (512 B, one mbarrier per slot) into the free slot while the octet
// works on the current one, whose slot -- once its coefficients are in registers -- is reused as the S<->C
// transposition tile.  Slots are kPolyWords = 152 words apart, i.e. 24 banks: the raw S-layout reads
// (word 8s + v) of the four octets of a warp fall into 32 distinct banks.  The global loads of the plain kernel
// (16 dependent-latency LDG per lane and iteration) become shared-memory loads behind an mbarrier wait that has
// normally completed an iteration earlier.
__global__ void __launch_bounds__(kThreads, 6) ntt_fwd_tma_kernel(uint32_t* __restrict__ polys, size_t n,
                                                                   const TwPair* __restrict__ tw) {
  __shared__ __align__(16) uint32_t slots[kOctetsPerCta * 2 * kPolyWords];
  __shared__ __align__(16) TwPair tws[128];
  __shared__ __align__(8) uint64_t bar, bars[kOctetsPerCta * 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int oct = lane >> 3, v = lane & 7, ob = warp * 4 + oct;
  uint32_t* slot0 = slots + ob * 2 * kPolyWords;
  uint64_t* obar = bars + ob * 2;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  if (v == 0) {
    mbar_init(obar, 1);
    mbar_init(obar + 1, 1);
  }
  fence_barrier_init();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, sizeof(tws));
    bulk_g2s(tws, tw, sizeof(tws), &bar);
  }
  const size_t stride = (size_t)gridDim.x * kOctetsPerCta, first = ((size_t)blockIdx.x * 4 + warp) * 4;
  const size_t n_it = first < n ? (n - first + stride - 1) / stride : 0;  // the same for the four octets of a warp
  auto poly_of = [&](size_t it) {
    const size_t p = first + it * stride + oct;
    return polys + (p < n ? p : n - 1) * (N / 2);  // idle octets recompute the last polynomial, never store
  };
  auto issue = [&](size_t it) {  // lane 0 of the octet
    uint64_t* b = obar + (it & 1);
    mbar_expect_tx(b, N * 2);
    bulk_g2s(slot0 + (it & 1) * kPolyWords, poly_of(it), N * 2, b);
  };
  if (v == 0) {
    if (n_it > 0) issue(0);
    if (n_it > 1) issue(1);
  }
  mbar_wait(&bar, 0);
  LaneTw t;
  load_lane_tw(t, tws, v);
  for (size_t it = 0; it < n_it; it++) {
    uint32_t* slot = slot0 + (it & 1) * kPolyWords;
    const bool active = first + it * stride + oct < n;
    mbar_wait(obar + (it & 1), (uint32_t)((it >> 1) & 1));
    int32_t r[32];
#pragma unroll
    for (int s = 0; s < 16; s++) unpack2_ct(slot[8 * s + v], r[2 * s], r[2 * s + 1]);
    __syncwarp();  // the raw polynomial is in registers: the slot becomes the transposition tile
    fwd_pass_S(r);
    store_S(slot, v, r);
    __syncwarp();
    load_C_ct(slot, v, r);
    __syncwarp();  // every lane is done with the tile: it may be refilled
    if (v == 0 && it + 2 < n_it) {
      fence_proxy_async();  // order the generic-proxy accesses above before the asynchronous write into the slot
      issue(it + 2);
    }
    fwd_pass_C(r, t);
    if (active) gstore_C(poly_of(it), v, r);
  }
}

// ---------------------------------------------------------------- fast-path kernels (see kyber.cuh, "low" format)
__device__ __forceinline__ void gstore_C_lo(uint32_t* __restrict__ poly, int v, const int32_t (&r)[32]) {
#pragma unroll
  for (int c = 0; c < 4; c++)
    stg_stream128(poly + 16 * v + 4 * c,
                  make_uint4(pack2_lo(r[8 * c], r[8 * c + 1]), pack2_lo(r[8 * c + 2], r[8 * c + 3]),
                             pack2_lo(r[8 * c + 4], r[8 * c + 5]), pack2_lo(r[8 * c + 6], r[8 * c + 7])));
}
__device__ __forceinline__ void gstore_S_lo(uint32_t* __restrict__ poly, int v, const int32_t (&r)[32]) {
#pragma unroll
  for (int s = 0; s < 16; s++) poly[8 * s + v] = pack2_lo(r[2 * s], r[2 * s + 1]);
}

constexpr int kSlotWords = 136;  // 512 B of coefficients + 32 B: consecutive octets start 8 banks apart

// Forward NTT.  Input staged by the TMA engine as in ntt_fwd_tma_kernel (SLOTS buffers per octet, refilled as soon as
// the raw words are in registers -- the transposition has its own tile here); in-contract polynomials run the
// 6-instruction butterflies on low-format registers with the 128-bit transposition, anything else the general code.
template <int SLOTS, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) ntt_fwd_fast_kernel(uint32_t* __restrict__ polys, size_t n,
                                                                       const TwPair* __restrict__ tw) {
  extern __shared__ __align__(16) unsigned char dsm[];
  TwPair* tws = reinterpret_cast<TwPair*>(dsm);                       // 128 x {zeta, zetaq} then 128 x {zp, kk}
  uint32_t* slots = reinterpret_cast<uint32_t*>(dsm + 2048);          // [SLOTS][16 octets][kSlotWords]
  unsigned char* tiles = dsm + 2048 + SLOTS * kOctetsPerCta * kSlotWords * 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tiles + kOctetsPerCta * kWideTileBytes);  // [16 octets][SLOTS], then 1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int oct = lane >> 3, v = lane & 7, ob = warp * 4 + oct;
  unsigned char* tile = tiles + ob * kWideTileBytes;
  uint64_t* obar = bars + ob * SLOTS;
  uint64_t* tbar = bars + kOctetsPerCta * SLOTS;
  if (threadIdx.x == 0) mbar_init(tbar, 1);
  if (v == 0)
    for (int k = 0; k < SLOTS; k++) mbar_init(obar + k, 1);
  fence_barrier_init();
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(tbar, 2048);
    bulk_g2s(tws, tw, 2048, tbar);
  }
  // Every octet walks its own polynomials p = first + oct + it * stride < n; the warp loops until its first octet (the
  // one with the most) is done, and an octet that has run out neither issues nor waits nor stores.
  const size_t stride = (size_t)gridDim.x * kOctetsPerCta, first = ((size_t)blockIdx.x * 4 + warp) * 4;
  const uint32_t n_it = first < n ? (uint32_t)((n - first + stride - 1) / stride) : 0;         // of the warp
  const uint32_t my_it = first + oct < n ? (uint32_t)((n - first - oct + stride - 1) / stride) : 0;  // of this octet
  const size_t step_words = stride * (N / 2);
  uint32_t* gp = polys + (first + oct) * (N / 2);   // the polynomial this octet transforms in iteration `it`
  const uint32_t* nxt = gp;                          // the next polynomial to request
  uint32_t issued = 0;
  auto issue = [&]() {  // lane 0 of the octet
    const uint32_t k = issued % SLOTS;
    mbar_expect_tx(obar + k, N * 2);
    bulk_g2s(slots + (k * kOctetsPerCta + ob) * kSlotWords, nxt, N * 2, obar + k);
    nxt += step_words;
    issued++;
  };
  if (v == 0)
    for (int k = 0; k < SLOTS; k++)
      if ((uint32_t)k < my_it) issue();
  mbar_wait(tbar, 0);
  LaneTwLow t;
  load_lane_tw_lo(t, reinterpret_cast<const TwLow*>(tws + 128), v);
  for (uint32_t it = 0; it < n_it; it++, gp += step_words) {
    const uint32_t k = it % SLOTS;
    const uint32_t* slot = slots + (k * kOctetsPerCta + ob) * kSlotWords;
    const bool active = it < my_it;
    uint32_t w[16];
    if (active) {
      mbar_wait(obar + k, (it / SLOTS) & 1);
#pragma unroll
      for (int s = 0; s < 16; s++) w[s] = slot[8 * s + v];
    } else {
#pragma unroll
      for (int s = 0; s < 16; s++) w[s] = 0;
    }
    const bool fast = __all_sync(0xffffffffu, words_in_range(w, kFwdBound));  // also: every lane has read its words
    if (v == 0 && issued < my_it) {
      fence_proxy_async();
      issue();
    }
    int32_t r[32];
    if (fast) {
#pragma unroll
      for (int s = 0; s < 16; s++) unpack2_lo(w[s], r[2 * s], r[2 * s + 1]);
      fwd_pass_S_lo(r);
      wide_store_S(tile, v, r);
      __syncwarp();
      wide_load_C(tile, v, r);
      fwd_pass_C_lo(r, t);
      if (active) gstore_C_lo(gp, v, r);
    } else {
#pragma unroll
      for (int s = 0; s < 16; s++) unpack2_ct(w[s], r[2 * s], r[2 * s + 1]);
      fwd_pass_S(r);
      store_S(reinterpret_cast<uint32_t*>(tile), v, r);
      __syncwarp();
      load_C_ct(reinterpret_cast<const uint32_t*>(tile), v, r);
      fwd_pass_C_smem(r, tws, v);
      if (active) gstore_C(gp, v, r);
    }
    __syncwarp();  // the tile is free again
  }
}
template <int SLOTS>
constexpr int fwd_fast_smem() {
  return 2048 + SLOTS * kOctetsPerCta * kSlotWords * 4 + kOctetsPerCta * kWideTileBytes + (kOctetsPerCta * SLOTS + 1) * 8;
}

// Inverse NTT: C-layout input as four 128-bit loads per lane (PREFETCH: the next polynomial's words are requested
// before the current one is transformed), the same fast / general split.
template <bool PREFETCH, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) ntt_inv_fast_kernel(uint32_t* __restrict__ polys, size_t n,
                                                                       const TwPair* __restrict__ tw) {
  __shared__ __align__(16) unsigned char tiles[kOctetsPerCta * kWideTileBytes];
  __shared__ __align__(16) TwPair tws[256];
  __shared__ __align__(8) uint64_t bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int oct = lane >> 3, v = lane & 7;
  unsigned char* tile = tiles + (warp * 4 + oct) * kWideTileBytes;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, sizeof(tws));
    bulk_g2s(tws, tw, sizeof(tws), &bar);
  }
  const size_t stride = (size_t)gridDim.x * kOctetsPerCta;
  auto poly_at = [&](size_t base) {
    const size_t p = base + oct;
    return polys + (p < n ? p : n - 1) * (N / 2);
  };
  auto fetch = [&](const uint32_t* poly, uint32_t (&w)[16]) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const uint4 q4 = ldg_stream128(poly + 16 * v + 4 * c);
      w[4 * c] = q4.x;
      w[4 * c + 1] = q4.y;
      w[4 * c + 2] = q4.z;
      w[4 * c + 3] = q4.w;
    }
  };
  size_t base = ((size_t)blockIdx.x * 4 + warp) * 4;
  uint32_t wn[16];
  if (PREFETCH && base < n) fetch(poly_at(base), wn);
  mbar_wait(&bar, 0);
  const volatile TwPair* tab = tws;
  const volatile TwLow* tabl = reinterpret_cast<const volatile TwLow*>(tws + 128);
  for (; base < n; base += stride) {
    const bool active = base + oct < n;
    uint32_t* poly = poly_at(base);
    uint32_t w[16];
    if (PREFETCH) {
#pragma unroll
      for (int i = 0; i < 16; i++) w[i] = wn[i];
      if (base + stride < n) fetch(poly_at(base + stride), wn);
    } else {
      fetch(poly, w);
    }
    const bool fast = __all_sync(0xffffffffu, words_in_range(w, kInvBound));
    int32_t r[32];
    if (fast) {
#pragma unroll
      for (int j = 0; j < 16; j++) unpack2_lo(w[j], r[2 * j], r[2 * j + 1]);
      inv_pass_C_lo(r, tabl, v);
      wide_store_C(tile, v, r);
      __syncwarp();
      wide_load_S(tile, v, r);
      inv_pass_S_lo(r, v);
      if (active) gstore_S_lo(poly, v, r);
    } else {
#pragma unroll
      for (int j = 0; j < 16; j++) unpack2(w[j], r[2 * j], r[2 * j + 1]);
      inv_pass_C_smem(r, tab, v);
      store_C(reinterpret_cast<uint32_t*>(tile), v, r);
      __syncwarp();
      load_S(reinterpret_cast<const uint32_t*>(tile), v, r);
      inv_pass_S(r, v);
      if (active) gstore_S(poly, v, r);
    }
    __syncwarp();
  }
}

static int grid_for(size_t units, int per_cta, int ctas_per_sm) {
  size_t want = (units + per_cta - 1) / per_cta;
  size_t cap = (size_t)kNumSM * ctas_per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace kyber

// ---------------------------------------------------------------- launchers (device pointers)
static int ntt_variant(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
int launch_kyber_ntt(int16_t* d_polys, size_t n, int inverse, const void* tw, cudaStream_t st) {
  using namespace kyber;
  if (n == 0) return 0;
  uint32_t* p = (uint32_t*)d_polys;
  const TwPair* t = (const TwPair*)tw;
  KernelScope ks(inverse ? KID_KYBER_INVNTT : KID_KYBER_NTT, st);
  if (inverse) {
    const int var = ntt_variant("CB200_INVNTT_VARIANT", 1);
    switch (var) {
      case 0: ntt_kernel<true><<<grid_for(n, kOctetsPerCta, 8), kThreads, 0, st>>>(p, n, t); break;
      case 1: ntt_inv_fast_kernel<false, 8><<<grid_for(n, kOctetsPerCta, 8), kThreads, 0, st>>>(p, n, t); break;
      case 2: ntt_inv_fast_kernel<false, 6><<<grid_for(n, kOctetsPerCta, 6), kThreads, 0, st>>>(p, n, t); break;
      default: ntt_inv_fast_kernel<true, 6><<<grid_for(n, kOctetsPerCta, 6), kThreads, 0, st>>>(p, n, t); break;
    }
  } else {
    const int var = ntt_variant("CB200_NTT_VARIANT", 1);
    switch (var) {
      case 0: ntt_fwd_tma_kernel<<<grid_for(n, kOctetsPerCta, 6), kThreads, 0, st>>>(p, n, t); break;
      case 1: ntt_fwd_fast_kernel<2, 5><<<grid_for(n, kOctetsPerCta, 5), kThreads, fwd_fast_smem<2>(), st>>>(p, n, t); break;
      case 2: ntt_fwd_fast_kernel<1, 6><<<grid_for(n, kOctetsPerCta, 6), kThreads, fwd_fast_smem<1>(), st>>>(p, n, t); break;
      default: ntt_fwd_fast_kernel<1, 5><<<grid_for(n, kOctetsPerCta, 5), kThreads, fwd_fast_smem<1>(), st>>>(p, n, t); break;
    }
  }
  CB200_CUDA(cudaGetLastError());
  return 0;
}

int launch_kyber_dot(int16_t* d_out, const int16_t* d_a, const int16_t* d_b, int k, size_t n, const void* tw,
                     cudaStream_t st) {
  using namespace kyber;
  if (n == 0) return 0;
  int grid = grid_for(n, kOctetsPerCta, 8);
  KernelScope ks(KID_KYBER_DOT, st);
  dot_kernel<<<grid, kThreads, 0, st>>>((uint32_t*)d_out, (const uint32_t*)d_a, (const uint32_t*)d_b, k, n,
                                        (const TwPair*)tw);
  CB200_CUDA(cudaGetLastError());
  return 0;
}

int launch_kyber_poly_op(int op, int16_t* d_out, const int16_t* d_a, const int16_t* d_b, size_t n, cudaStream_t st) {
  using namespace kyber;
  if (n == 0) return 0;
  size_t nvec = n * (N / 8);
  int grid = grid_for(nvec, 256, 8);
  uint4* o = (uint4*)d_out;
  const uint4 *a = (const uint4*)d_a, *b = (const uint4*)d_b;
  KernelScope ks(KID_KYBER_EW, st);
  switch (op) {
    case OP_ADD: ew_kernel<OP_ADD><<<grid, 256, 0, st>>>(o, a, b, nvec); break;
    case OP_SUB: ew_kernel<OP_SUB><<<grid, 256, 0, st>>>(o, a, b, nvec); break;
    case OP_BARRETT: ew_kernel<OP_BARRETT><<<grid, 256, 0, st>>>(o, a, b, nvec); break;
    case OP_NORMALIZE: ew_kernel<OP_NORMALIZE><<<grid, 256, 0, st>>>(o, a, b, nvec); break;
    case OP_TOMONT: ew_kernel<OP_TOMONT><<<grid, 256, 0, st>>>(o, a, b, nvec); break;
    default: set_error("cb200_kyber_poly_op: unknown op %d", op); return -1;
  }
  CB200_CUDA(cudaGetLastError());
  return 0;
}

// host-side twiddle table: {Zetas[k], (Zetas[k]*q^-1 mod 2^16) << 16}
void kyber_fill_twiddles(int32_t* out /* 128 x {zeta, zetaq}, then 128 x {zp, kk} */) {
  for (int i = 0; i < 128; i++) {
    out[2 * i] = kyber::zeta_of(i);
    out[2 * i + 1] = kyber::zetaq_of(i);
    out[256 + 2 * i] = kyber::zp_of(i);
    out[256 + 2 * i + 1] = kyber::kk_of(i);
  }
}

}  // namespace cb200
