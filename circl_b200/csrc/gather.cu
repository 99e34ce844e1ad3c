// circl_b200/csrc/gather.cu -- the result gather of the one-process-per-GPU launch (north_star:
// "gather ciphertexts/signatures back to rank 0").
//
// The path has no exchange during compute; what remains is moving every rank's rows into rank 0's buffer.  NCCL
// send/recv does that with kernels, which take SMs from an ALU-saturated compute stream (round 1: 3.3 ms exposed at
// N = 8).  Here rank 0 allocates the destination and publishes a CUDA IPC handle; every other rank maps it and pushes
// its rows with cudaMemcpyAsync into the mapped peer memory: NVLink traffic issued by the copy engines, no SM involved,
// on a copy stream that waits for the producing kernels through an event.  NCCL (torch.distributed) stays for the
// rendezvous -- the 64-byte handle travels through it -- and for barriers.
#include <string.h>

#include "../../include/circl_b200.h"
#include "common.cuh"
#include "context.h"

using namespace cb200;

extern "C" {

int cb200_gather_alloc(size_t bytes, void** dev_ptr, uint8_t* handle) {
  int rc = require_ready();
  if (rc) return rc;
  if (!dev_ptr || !handle || bytes == 0) {
    set_error("cb200_gather_alloc: bad argument");
    return CB200_ERR_ARG;
  }
  void* p = nullptr;
  CB200_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cb200_gather_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return -100 - (int)e;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == CB200_GATHER_HANDLE_BYTES, "handle size");
  memcpy(handle, &h, sizeof h);
  *dev_ptr = p;
  return 0;
}

int cb200_gather_free(void* dev_ptr) {
  if (dev_ptr) CB200_CUDA(cudaFree(dev_ptr));
  return 0;
}

int cb200_gather_open(const uint8_t* handle, void** peer_ptr) {
  int rc = require_ready();
  if (rc) return rc;
  if (!handle || !peer_ptr) {
    set_error("cb200_gather_open: bad argument");
    return CB200_ERR_ARG;
  }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  CB200_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int cb200_gather_close(void* peer_ptr) {
  if (peer_ptr) CB200_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

int cb200_gather_push(void* dst, const void* src, size_t bytes) {
  int rc = require_ready();
  if (rc) return rc;
  if (bytes == 0) return 0;
  if (!dst || !src) {
    set_error("cb200_gather_push: null pointer");
    return CB200_ERR_ARG;
  }
  DeviceCall call(src);  // the work set of (GPU that holds the rows, this thread's stream)
  if (call.rc) return call.rc;
  CB200_CUDA(cudaEventRecord(call.ws->ev_copy, call.st));
  CB200_CUDA(cudaStreamWaitEvent(call.ws->copy, call.ws->ev_copy, 0));
  CB200_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, call.ws->copy));
  return 0;
}

int cb200_gather_flush(const void* any_local_row, int wait_on_host) {
  int rc = require_ready();
  if (rc) return rc;
  DeviceCall call(any_local_row);
  if (call.rc) return call.rc;
  if (wait_on_host) {
    CB200_CUDA(cudaStreamSynchronize(call.ws->copy));
  } else {
    CB200_CUDA(cudaEventRecord(call.ws->ev_copy, call.ws->copy));
    CB200_CUDA(cudaStreamWaitEvent(call.st, call.ws->ev_copy, 0));
  }
  return 0;
}

}  // extern "C"
