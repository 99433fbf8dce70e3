// Shared helpers for the tsu_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tsu_b200.h"

#define TSU_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return TSU_ERR_INVALID_ARG; \
  } while (0)

// Kernel launches never throw: the C-ABI returns the cudaError_t (positive) of the launch.
#define TSU_RETURN_LAUNCH_STATUS()          \
  do {                                      \
    cudaError_t e__ = cudaGetLastError();   \
    return e__ == cudaSuccess ? TSU_OK : (int)e__; \
  } while (0)

static inline cudaStream_t tsu_stream(uintptr_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define TSU_NUM_SMS 148  // B200: 2 dies x 74 SMs

__device__ __forceinline__ uint32_t tsu_lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
  return (a & b) | (a & c) | (b & c);
}

// warp-wide sum of a 64-bit count
__device__ __forceinline__ unsigned long long tsu_warp_sum(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
