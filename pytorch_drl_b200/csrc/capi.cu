// Library-wide C-ABI entry points: versioning, status strings, device check.
#include "common.cuh"

namespace drl {
thread_local cudaError_t g_last_cuda_error = cudaSuccess;

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace drl

extern "C" int drl_abi_version(void) { return DRL_B200_ABI_VERSION; }

extern "C" const char* drl_status_string(int s) {
  switch (s) {
    case DRL_OK: return "ok";
    case DRL_ERR_BAD_SHAPE: return "bad shape";
    case DRL_ERR_BAD_ARG: return "bad argument";
    case DRL_ERR_MISALIGNED: return "misaligned pointer";
    case DRL_ERR_CUDA: return "CUDA error";
    case DRL_ERR_NO_DEVICE: return "no sm_100 CUDA device (there is no CPU fallback)";
    case DRL_ERR_UNSUPPORTED: return "unsupported";
    case DRL_ERR_NCCL: return "NCCL error";
    default: return "unknown status";
  }
}

extern "C" const char* drl_last_cuda_error(void) {
  return drl::g_last_cuda_error == cudaSuccess ? "" : cudaGetErrorString(drl::g_last_cuda_error);
}

extern "C" int drl_check_device(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) return DRL_ERR_NO_DEVICE;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, device) != cudaSuccess) return DRL_ERR_NO_DEVICE;
  return p.major == 10 ? DRL_OK : DRL_ERR_NO_DEVICE;
}

extern "C" int drl_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return DRL_ERR_BAD_ARG;
  DRL_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, drl::as_stream(stream)));
  return DRL_OK;
}

// Copy by a kernel instead of a copy engine: `src` may be PINNED HOST memory (unified addressing: the SMs read it over PCIe).  For
// small host-to-device transfers that must not queue behind bulk uploads on the host-to-device copy engine (the per-epoch
// minibatch indices of PPO.optimize while the next rollout is being uploaded).
namespace drl {
__global__ void __launch_bounds__(256) copy_words_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
}  // namespace drl

extern "C" int drl_copy_by_kernel(void* dst, const void* src, size_t bytes, void* stream) {
  if (!dst || !src) return DRL_ERR_BAD_ARG;
  if ((bytes & 3) || (reinterpret_cast<uintptr_t>(dst) & 3) || (reinterpret_cast<uintptr_t>(src) & 3)) return DRL_ERR_MISALIGNED;
  if (bytes == 0) return DRL_OK;
  const int64_t n = (int64_t)(bytes / 4);
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 296) blocks = 296;
  drl::copy_words_kernel<<<blocks, 256, 0, drl::as_stream(stream)>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst), n);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
