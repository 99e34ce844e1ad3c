// circl_b200/csrc/context.h -- process-wide state behind the C ABI (one process per GPU).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>
#include <functional>
#include <mutex>
#include <vector>

#include "launch.h"

namespace cb200 {

struct ProfRec {
  int id;
  cudaEvent_t a, b;
};

struct Ctx {
  bool ready = false;
  int device = -1;
  int sm_count = 148;           // multiprocessors of the device (sizes the persistent grids)
  cudaStream_t own = nullptr;   // library stream
  cudaStream_t cur = nullptr;   // stream used for device-pointer calls (own or user supplied)
  cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};  // host-pointer calls: H2D -> kernels -> D2H per chunk
  void* kyber_tw = nullptr;     // 128 x {zeta, zetaq}
  void* dil_tw = nullptr;       // 256 x {zeta, invzeta}
  void* small = nullptr;        // 256-byte device buffer (ML-DSA context string)
  void* x25519_table = nullptr; // 32 KiB: multiples 1..8 of 256^i B for the fixed-base X25519 KeyGen (x25519.cuh)
  std::atomic<uint64_t> launches{0};
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::mutex prof_mu;
  std::mutex mu;                // serialises host-pointer calls (device scratch is shared)
  // grow-only device scratch, one per pipeline slot
  void* scratch[3] = {nullptr, nullptr, nullptr};
  size_t scratch_bytes[3] = {0, 0, 0};
  // grow-only work areas for the multi-kernel pipelines (ML-KEM / ML-DSA)
  // slot 0..2 belong to the staging pipeline streams, slot 3 to device-pointer calls
  // two internal "lanes" per work slot: consecutive sub-batches of a pipeline alternate between them so that
  // the tail of one sub-batch's kernels overlaps the head of the next (fork/join on events around them)
  cudaStream_t lane[4][2] = {};
  cudaEvent_t ev_fork[4] = {}, ev_join[4][2] = {};
  void* pinned = nullptr;       // grow-only pinned host staging for small per-op outputs (status bytes)
  size_t pinned_bytes = 0;
  // work[0..3]: the lattice flows of pipeline slots 0..2 and of device-pointer calls (3);
  // work[4..7]: second level, for the flows of hybrid.cu that call into the lattice flows of the same slot
  void* work[8] = {};
  size_t work_bytes[8] = {};
};

// Kernel classes for the optional per-kernel event timing (cb200_profile_*).
enum KernelId {
  KID_KYBER_NTT = 0, KID_KYBER_INVNTT, KID_KYBER_DOT, KID_KYBER_EW,
  KID_MLKEM_HASH_EK, KID_MLKEM_G, KID_MLKEM_SAMPLE, KID_MLKEM_ENCRYPT,
  KID_DIL_NTT, KID_DIL_INVNTT, KID_DIL_DOT, KID_DIL_EW,
  KID_MLDSA_EXPAND, KID_MLDSA_MU, KID_MLDSA_MASK, KID_MLDSA_W, KID_MLDSA_CHALLENGE, KID_MLDSA_RESPONSE,
  KID_MLDSA_COMPACT, KID_X25519, KID_HYBRID_GLUE, KID_COUNT
};
const char* kernel_name(int id);

Ctx& ctx();
// Brackets one kernel launch with events when profiling is on; always counts the launch.
struct KernelScope {
  cudaStream_t st;
  int id;
  cudaEvent_t a = nullptr;
  KernelScope(int id, cudaStream_t st);
  ~KernelScope();
};
int require_ready();
bool is_device_ptr(const void* p);
int ensure_scratch(int slot, size_t bytes);
int ensure_work(int slot, size_t bytes, void** out);
int ensure_pinned(size_t bytes, void** out);
inline void count_launch(uint64_t n = 1) { ctx().launches.fetch_add(n, std::memory_order_relaxed); }

// One buffer of a batched call: `unit` bytes per batch element; stride 0 = shared by all elements.
struct Buf {
  const void* host_in = nullptr;  // copied host -> device before the kernels (may be null)
  void* host_out = nullptr;       // copied device -> host after the kernels (may be null)
  size_t unit = 0;
  bool shared = false;            // whole buffer is `unit` bytes, identical for every element
  size_t host_stride = 0;         // bytes between elements in host memory (0 = dense = unit)
};

// Runs `body(dev_ptrs, count, stream)` over the batch in chunks, overlapping the
// host<->device copies of one chunk with the kernels of another on three streams.
// dev_ptrs[i] is the device image of bufs[i] for the current chunk.
int run_staged(std::vector<Buf>& bufs, size_t n, size_t chunk,
               const std::function<int(void** dev, size_t count, size_t first, cudaStream_t st, int slot)>& body);

}  // namespace cb200
