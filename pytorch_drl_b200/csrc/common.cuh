// Shared helpers for the drl_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/drl_b200.h"
#include "profile.cuh"

namespace drl {

extern thread_local cudaError_t g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = e;
  return DRL_ERR_CUDA;
}

#define DRL_CUDA(expr)                                   \
  do {                                                   \
    cudaError_t _e = (expr);                             \
    if (_e != cudaSuccess) return ::drl::cuda_fail(_e);  \
  } while (0)

#define DRL_LAUNCH_CHECK()                               \
  do {                                                   \
    ::drl::count_launch();                               \
    cudaError_t _e = cudaGetLastError();                 \
    if (_e != cudaSuccess) return ::drl::cuda_fail(_e);  \
  } while (0)

#define DRL_TRY(expr)                  \
  do {                                 \
    int _s = (expr);                   \
    if (_s != DRL_OK) return _s;       \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (148 on B200); cached per process.
int sm_count();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace drl
