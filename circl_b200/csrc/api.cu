// circl_b200/csrc/api.cu -- C ABI of libcirclb200.so (declared in include/circl_b200.h):
// context, pointer classification, host staging pipeline and the Kyber ring entry points.
#include <stdarg.h>
#include <string.h>

#include "../../include/circl_b200.h"
#include "common.cuh"
#include "context.h"

namespace cb200 {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

Ctx& ctx() {
  static Ctx c;
  return c;
}

static const char* kKernelNames[KID_COUNT] = {
    "kyber_ntt", "kyber_invntt", "kyber_dot", "kyber_elementwise",
    "mlkem_hash_ek", "mlkem_g", "mlkem_sample", "mlkem_encrypt",
    "dil_ntt", "dil_invntt", "dil_dot", "dil_elementwise",
    "mldsa_expand_key", "mldsa_mu_rhoprime", "mldsa_mask", "mldsa_w", "mldsa_challenge", "mldsa_response",
    "mldsa_compact", "x25519", "hybrid_glue"};
const char* kernel_name(int id) { return (id >= 0 && id < KID_COUNT) ? kKernelNames[id] : "?"; }

KernelScope::KernelScope(int id_, cudaStream_t st_) : st(st_), id(id_) {
  Ctx& c = ctx();
  c.launches.fetch_add(1, std::memory_order_relaxed);
  if (c.profiling) {
    cudaEventCreate(&a);
    cudaEventRecord(a, st);
  }
}
KernelScope::~KernelScope() {
  if (!a) return;
  Ctx& c = ctx();
  cudaEvent_t b;
  cudaEventCreate(&b);
  cudaEventRecord(b, st);
  std::lock_guard<std::mutex> lock(c.prof_mu);
  c.prof.push_back(ProfRec{id, a, b});
}

int require_ready() {
  if (!ctx().ready) {
    set_error("cb200: not initialised (call cb200_init; there is no CPU fallback)");
    return CB200_ERR_NOT_INIT;
  }
  return 0;
}

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int ensure_scratch(int slot, size_t bytes) {
  Ctx& c = ctx();
  if (c.scratch_bytes[slot] >= bytes) return 0;
  if (c.scratch[slot]) CB200_CUDA(cudaFree(c.scratch[slot]));
  c.scratch[slot] = nullptr;
  c.scratch_bytes[slot] = 0;
  CB200_CUDA(cudaMalloc(&c.scratch[slot], bytes));
  c.scratch_bytes[slot] = bytes;
  return 0;
}

int ensure_work(int slot, size_t bytes, void** out) {
  Ctx& c = ctx();
  if (c.work_bytes[slot] < bytes) {
    CB200_CUDA(cudaDeviceSynchronize());
    if (c.work[slot]) CB200_CUDA(cudaFree(c.work[slot]));
    c.work[slot] = nullptr;
    c.work_bytes[slot] = 0;
    CB200_CUDA(cudaMalloc(&c.work[slot], bytes));
    c.work_bytes[slot] = bytes;
  }
  *out = c.work[slot];
  return 0;
}

int ensure_pinned(size_t bytes, void** out) {
  Ctx& c = ctx();
  if (c.pinned_bytes < bytes) {
    if (c.pinned) CB200_CUDA(cudaFreeHost(c.pinned));
    c.pinned = nullptr;
    c.pinned_bytes = 0;
    CB200_CUDA(cudaHostAlloc(&c.pinned, bytes, cudaHostAllocDefault));
    c.pinned_bytes = bytes;
  }
  *out = c.pinned;
  return 0;
}

int run_staged(std::vector<Buf>& bufs, size_t n, size_t chunk,
               const std::function<int(void** dev, size_t count, size_t first, cudaStream_t st, int slot)>& body) {
  Ctx& c = ctx();
  std::lock_guard<std::mutex> lock(c.mu);
  if (chunk == 0 || chunk > n) chunk = n;
  const size_t nb = bufs.size();
  std::vector<size_t> off(nb + 1, 0);
  for (size_t i = 0; i < nb; i++) {
    size_t sz = bufs[i].shared ? bufs[i].unit : bufs[i].unit * chunk;
    off[i + 1] = off[i] + ((sz + 255) & ~(size_t)255);
  }
  for (int s = 0; s < 3; s++) {
    int rc = ensure_scratch(s, off[nb]);
    if (rc) return rc;
  }
  // Chunk schedule: the first and the last chunks are short (1/4, 1/4, 1/2 of `chunk` on the way up, the mirror
  // image on the way down) so that the part of the pipeline that cannot overlap -- the first input copy and the last
  // kernels + output copy -- is a quarter of a chunk instead of a whole one.  Only worth it for long batches.
  std::vector<size_t> sched;
  {
    const size_t ramp[3] = {chunk / 4, chunk / 4, chunk / 2};
    size_t left = n;
    std::vector<size_t> tail;
    if (n >= 6 * chunk && chunk >= 4096) {
      for (size_t r : ramp) {
        sched.push_back(r);
        tail.push_back(r);
        left -= 2 * r;
      }
    }
    while (left > 0) {
      const size_t cnt = left < chunk ? left : chunk;
      sched.push_back(cnt);
      left -= cnt;
    }
    for (size_t k = tail.size(); k-- > 0;) sched.push_back(tail[k]);
  }
  std::vector<void*> dev(nb);
  int slot = 0, rc = 0;
  size_t first = 0;
  for (size_t ci = 0; ci < sched.size() && rc == 0; first += sched[ci], ci++, slot = (slot + 1) % 3) {
    const size_t count = sched[ci];
    cudaStream_t st = c.pipe[slot];
    for (size_t i = 0; i < nb; i++) {
      dev[i] = (char*)c.scratch[slot] + off[i];
      if (bufs[i].host_in) {
        const size_t hs = bufs[i].host_stride ? bufs[i].host_stride : bufs[i].unit;
        const char* src = (const char*)bufs[i].host_in + (bufs[i].shared ? 0 : first * hs);
        if (bufs[i].shared || hs == bufs[i].unit) {
          size_t sz = bufs[i].shared ? bufs[i].unit : count * bufs[i].unit;
          CB200_CUDA(cudaMemcpyAsync(dev[i], src, sz, cudaMemcpyHostToDevice, st));
        } else {
          CB200_CUDA(cudaMemcpy2DAsync(dev[i], bufs[i].unit, src, hs, bufs[i].unit, count, cudaMemcpyHostToDevice, st));
        }
      }
    }
    rc = body(dev.data(), count, first, st, slot);
    if (rc) break;
    for (size_t i = 0; i < nb; i++)
      if (bufs[i].host_out && !bufs[i].shared)
        CB200_CUDA(cudaMemcpyAsync((char*)bufs[i].host_out + first * bufs[i].unit, dev[i], count * bufs[i].unit,
                                   cudaMemcpyDeviceToHost, st));
  }
  for (int s = 0; s < 3; s++) CB200_CUDA(cudaStreamSynchronize(c.pipe[s]));
  return rc;
}

// number of polynomials per staging chunk: 64 MiB of int16 polys keeps three
// chunks in flight well under any memory pressure and amortises launch latency.
static constexpr size_t kPolyChunk = 1u << 17;

}  // namespace cb200

using namespace cb200;

extern "C" {

const char* cb200_version(void) { return "circl_b200 0.1 (sm_100a)"; }
const char* cb200_last_error(void) { return g_err; }

int cb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int cb200_init(int device) {
  Ctx& c = ctx();
  if (c.ready && c.device == device) return 0;
  if (c.ready) cb200_shutdown();
  int n = cb200_device_count();
  if (n <= 0) {
    set_error("cb200_init: no CUDA device visible (this library has no CPU fallback)");
    return CB200_ERR_NOT_INIT;
  }
  if (device < 0 || device >= n) {
    set_error("cb200_init: device %d out of range (have %d)", device, n);
    return CB200_ERR_ARG;
  }
  CB200_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CB200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("cb200_init: device %d is sm_%d%d; this build contains only sm_100a code", device, prop.major, prop.minor);
    return CB200_ERR_NOT_INIT;
  }
  c.sm_count = prop.multiProcessorCount;
  CB200_CUDA(cudaStreamCreateWithFlags(&c.own, cudaStreamNonBlocking));
  for (int s = 0; s < 3; s++) CB200_CUDA(cudaStreamCreateWithFlags(&c.pipe[s], cudaStreamNonBlocking));
  for (int w = 0; w < 4; w++) {
    CB200_CUDA(cudaEventCreateWithFlags(&c.ev_fork[w], cudaEventDisableTiming));
    for (int l = 0; l < 2; l++) {
      CB200_CUDA(cudaStreamCreateWithFlags(&c.lane[w][l], cudaStreamNonBlocking));
      CB200_CUDA(cudaEventCreateWithFlags(&c.ev_join[w][l], cudaEventDisableTiming));
    }
  }
  c.cur = nullptr;  // CUDA legacy default stream until the caller names one
  int32_t ktw[256];
  kyber_fill_twiddles(ktw);
  CB200_CUDA(cudaMalloc(&c.kyber_tw, sizeof ktw));
  CB200_CUDA(cudaMemcpy(c.kyber_tw, ktw, sizeof ktw, cudaMemcpyHostToDevice));
  int rc = init_extra_tables();
  if (rc) return rc;
  c.device = device;
  c.launches = 0;
  c.ready = true;
  return 0;
}

void cb200_shutdown(void) {
  Ctx& c = ctx();
  if (!c.ready) return;
  cudaDeviceSynchronize();
  for (int s = 0; s < 3; s++) {
    if (c.scratch[s]) cudaFree(c.scratch[s]);
    c.scratch[s] = nullptr;
    c.scratch_bytes[s] = 0;
    if (c.pipe[s]) cudaStreamDestroy(c.pipe[s]);
    c.pipe[s] = nullptr;
  }
  for (int w = 0; w < 4; w++) {
    if (c.ev_fork[w]) cudaEventDestroy(c.ev_fork[w]);
    c.ev_fork[w] = nullptr;
    for (int l = 0; l < 2; l++) {
      if (c.lane[w][l]) cudaStreamDestroy(c.lane[w][l]);
      if (c.ev_join[w][l]) cudaEventDestroy(c.ev_join[w][l]);
      c.lane[w][l] = nullptr;
      c.ev_join[w][l] = nullptr;
    }
  }
  for (int s = 0; s < 8; s++) {
    if (c.work[s]) cudaFree(c.work[s]);
    c.work[s] = nullptr;
    c.work_bytes[s] = 0;
  }
  if (c.pinned) cudaFreeHost(c.pinned);
  c.pinned = nullptr;
  c.pinned_bytes = 0;
  if (c.kyber_tw) cudaFree(c.kyber_tw);
  if (c.dil_tw) cudaFree(c.dil_tw);
  if (c.small) cudaFree(c.small);
  if (c.x25519_table) cudaFree(c.x25519_table);
  c.kyber_tw = c.dil_tw = c.small = c.x25519_table = nullptr;
  if (c.own) cudaStreamDestroy(c.own);
  c.own = c.cur = nullptr;
  c.ready = false;
  c.device = -1;
}

int cb200_set_stream(void* s) {
  int rc = require_ready();
  if (rc) return rc;
  ctx().cur = (cudaStream_t)s;
  return 0;
}

int cb200_synchronize(void) {
  int rc = require_ready();
  if (rc) return rc;
  CB200_CUDA(cudaStreamSynchronize(ctx().cur));
  return 0;
}

void* cb200_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    set_error("cb200_host_alloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
void cb200_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

uint64_t cb200_launch_count(void) { return ctx().launches.load(); }

int cb200_profile_enable(int on) {
  int rc = require_ready();
  if (rc) return rc;
  ctx().profiling = on != 0;
  return 0;
}
int cb200_profile_kernel_count(void) { return KID_COUNT; }
const char* cb200_profile_kernel_name(int id) { return kernel_name(id); }
int cb200_profile_read(double* ms_total, uint64_t* launches, int n) {
  int rc = require_ready();
  if (rc) return rc;
  Ctx& c = ctx();
  CB200_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lock(c.prof_mu);
  for (int i = 0; i < n; i++) {
    ms_total[i] = 0;
    launches[i] = 0;
  }
  for (ProfRec& r : c.prof) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.id < n) {
      ms_total[r.id] += ms;
      launches[r.id] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  c.prof.clear();
  return 0;
}

// ---------------------------------------------------------------- Kyber ring ops
int cb200_kyber_ntt(int16_t* polys, size_t n, int inverse) {
  int rc = require_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  if (!polys) {
    set_error("cb200_kyber_ntt: null pointer");
    return CB200_ERR_ARG;
  }
  if (is_device_ptr(polys)) return launch_kyber_ntt(polys, n, inverse, ctx().kyber_tw, ctx().cur);
  std::vector<Buf> bufs(1);
  bufs[0] = Buf{polys, polys, 512, false, 0};
  return run_staged(bufs, n, kPolyChunk, [&](void** d, size_t cnt, size_t, cudaStream_t st, int) {
    return launch_kyber_ntt((int16_t*)d[0], cnt, inverse, ctx().kyber_tw, st);
  });
}

int cb200_kyber_dot(int16_t* out, const int16_t* a, const int16_t* b, int k, size_t n) {
  int rc = require_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  if (!out || !a || !b || k < 1 || k > 8) {
    set_error("cb200_kyber_dot: bad argument");
    return CB200_ERR_ARG;
  }
  const bool dev = is_device_ptr(out);
  if (dev != is_device_ptr(a) || dev != is_device_ptr(b)) {
    set_error("cb200_kyber_dot: mixed host/device pointers");
    return CB200_ERR_ARG;
  }
  if (dev) return launch_kyber_dot(out, a, b, k, n, ctx().kyber_tw, ctx().cur);
  std::vector<Buf> bufs(3);
  bufs[0] = Buf{nullptr, out, 512, false, 0};
  bufs[1] = Buf{a, nullptr, 512 * (size_t)k, false, 0};
  bufs[2] = Buf{b, nullptr, 512 * (size_t)k, false, 0};
  return run_staged(bufs, n, kPolyChunk / k, [&](void** d, size_t cnt, size_t, cudaStream_t st, int) {
    return launch_kyber_dot((int16_t*)d[0], (const int16_t*)d[1], (const int16_t*)d[2], k, cnt, ctx().kyber_tw, st);
  });
}

int cb200_kyber_mulhat(int16_t* out, const int16_t* a, const int16_t* b, size_t n) {
  return cb200_kyber_dot(out, a, b, 1, n);
}

int cb200_kyber_poly_op(int op, int16_t* out, const int16_t* a, const int16_t* b, size_t n) {
  int rc = require_ready();
  if (rc) return rc;
  if (n == 0) return 0;
  const bool binary = (op == CB200_OP_ADD || op == CB200_OP_SUB);
  if (!out || !a || (binary && !b) || op < 0 || op > CB200_OP_TOMONT) {
    set_error("cb200_kyber_poly_op: bad argument");
    return CB200_ERR_ARG;
  }
  const bool dev = is_device_ptr(out);
  if (dev != is_device_ptr(a) || (binary && dev != is_device_ptr(b))) {
    set_error("cb200_kyber_poly_op: mixed host/device pointers");
    return CB200_ERR_ARG;
  }
  if (dev) return launch_kyber_poly_op(op, out, a, b, n, ctx().cur);
  std::vector<Buf> bufs(3);
  bufs[0] = Buf{nullptr, out, 512, false, 0};
  bufs[1] = Buf{a, nullptr, 512, false, 0};
  bufs[2] = Buf{binary ? b : nullptr, nullptr, 512, false, 0};
  return run_staged(bufs, n, kPolyChunk, [&](void** d, size_t cnt, size_t, cudaStream_t st, int) {
    return launch_kyber_poly_op(op, (int16_t*)d[0], (const int16_t*)d[1], (const int16_t*)d[2], cnt, st);
  });
}

}  // extern "C"
