// Device-level entry points of the lattice KEM flows (mlkem.cu) for the callers in hybrid.cu.
// All pointers are device pointers; `mlkem` selects FIPS 203 ML-KEM (1) or the round-3 Kyber FO wrapper (0);
// `slot` is the work-area slot (0..2 staging pipeline, 3 device-pointer calls).  Buffers and strides must be
// 16-byte aligned.  status may be null; round-3 Kyber never writes it.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace cb200 {
namespace mlkem {

int dev_keygen(int k, int mlkem, const uint8_t* seeds, uint8_t* ek, uint8_t* dk, size_t n, cudaStream_t st, int slot);
int dev_encaps(int k, int mlkem, const uint8_t* ek, size_t ek_stride, const uint8_t* seeds, uint8_t* ct, uint8_t* ss,
               uint8_t* status, size_t n, cudaStream_t st, int slot);
// one decapsulation key per op (dk_stride >= key size)
int dev_decaps(int k, int mlkem, const uint8_t* dk, size_t dk_stride, const uint8_t* ct, uint8_t* ss, uint8_t* status, size_t n,
               cudaStream_t st, int slot);

// dst[i*width .. ) = src[0 .. width) for i < n (width a multiple of 16, 16-byte aligned buffers)
int replicate_rows(const uint8_t* src, uint8_t* dst, size_t width, size_t n, cudaStream_t st);

}  // namespace mlkem
}  // namespace cb200
