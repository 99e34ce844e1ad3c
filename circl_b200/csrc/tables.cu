// circl_b200/csrc/tables.cu -- device tables built at cb200_init time.
#include "common.cuh"
#include "context.h"

namespace cb200 {
int init_extra_tables() {
  Ctx& c = ctx();
  uint32_t tw[512];
  dil_fill_twiddles(tw);
  if (c.dil_tw) cudaFree(c.dil_tw);
  CB200_CUDA(cudaMalloc(&c.dil_tw, sizeof tw));
  CB200_CUDA(cudaMemcpy(c.dil_tw, tw, sizeof tw, cudaMemcpyHostToDevice));
  if (!c.small) CB200_CUDA(cudaMalloc(&c.small, 256));
  return 0;
}
}  // namespace cb200
