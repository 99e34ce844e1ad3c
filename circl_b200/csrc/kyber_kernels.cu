// circl_b200/csrc/kyber_kernels.cu -- batched q=3329 kernels (sm_100a) and their launchers.
//
// cb200_kyber_ntt      <- (*Poly).NTT / InvNTT      pke/kyber/internal/common/generic.go:24,36 (stubs_amd64.go:8-14)
// cb200_kyber_mulhat   <- (*Poly).MulHat            generic.go:49 (stubs_amd64.go:17)
// cb200_kyber_dot      <- PolyDotHat                pke/kyber/kyber768/internal/vec.go:30-37
// cb200_kyber_poly_op  <- Add/Sub/BarrettReduce/Normalize/ToMont  generic.go:7-77, poly.go:48
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

static int grid_for(size_t units, int per_cta, int ctas_per_sm) {
  size_t want = (units + per_cta - 1) / per_cta;
  size_t cap = (size_t)kNumSM * ctas_per_sm;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace kyber

// ---------------------------------------------------------------- launchers (device pointers)
int launch_kyber_ntt(int16_t* d_polys, size_t n, int inverse, const void* tw, cudaStream_t st) {
  using namespace kyber;
  if (n == 0) return 0;
  int grid = grid_for(n, kOctetsPerCta, 8);
  KernelScope ks(inverse ? KID_KYBER_INVNTT : KID_KYBER_NTT, st);
  if (inverse)
    ntt_kernel<true><<<grid, kThreads, 0, st>>>((uint32_t*)d_polys, n, (const TwPair*)tw);
  else
    ntt_fwd_tma_kernel<<<grid_for(n, kOctetsPerCta, 6), kThreads, 0, st>>>((uint32_t*)d_polys, n, (const TwPair*)tw);
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
void kyber_fill_twiddles(int32_t* out /* 128 x 2 */) {
  for (int i = 0; i < 128; i++) {
    out[2 * i] = kyber::zeta_of(i);
    out[2 * i + 1] = kyber::zetaq_of(i);
  }
}

}  // namespace cb200
