// Micro-benchmark: cycles per tcgen05.mma (M=128, K=16, bf16) as a function of N, operand majorness,
// the number of independent accumulation chains, A-from-TMEM, and co-resident CTAs.  Timing only.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I pytorch_drl_b200/csrc -o /tmp/mma_rate scripts/microbench/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"

using namespace drl::umma;

template <int N, int A_MN, int B_MN, int CHAINS, int A_TMEM>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x * 16; i < 96 * 1024; i += blockDim.x * 16) st_shared_v4(smem_u32(smem + i), make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = slot;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc_f16(1u, 128, N, A_MN, B_MN);
    // A: 64 KB region (K-major: 128 rows x 128 B, k-step +32 B; MN-major: 2 atoms of 64 x 8-row groups, LBO = 8 KB
    // between the atoms, k-step of 16 rows = +2 KB).  B: 32 KB region after it.
    constexpr uint32_t a_hi = desc_hi(1024, LAYOUT_SW128);
    constexpr uint32_t b_hi = (B_MN && N < 64) ? desc_hi(512, LAYOUT_SW64) : desc_hi(1024, LAYOUT_SW128);
    const uint32_t a_lo0 = desc_lo(smem_u32(smem), A_MN ? 8192 : 0);
    const uint32_t b_lo0 = desc_lo(smem_u32(smem) + 65536, B_MN ? (N > 64 ? 8192 : 0) : 0);
    constexpr uint32_t a_step = A_MN ? (2048 >> 4) : 2, b_step = B_MN ? ((N < 64 ? 1024 : 2048) >> 4) : 2;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int ch = 0; ch < CHAINS; ++ch) {
          const uint32_t d = tb + ch * N;
          if (A_TMEM) mma_f16_ts<true>(d, tb + 256 + k * 8, b_lo0 + k * b_step, b_hi, idesc);
          else mma_f16_lohi<true>(d, a_lo0 + ch * (16384 >> 4) + k * a_step, a_hi, b_lo0 + k * b_step, b_hi, idesc);
        }
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

static long long* d_out;
template <int N, int A_MN, int B_MN, int CHAINS, int A_TMEM>
void run(int grid) {
  const int smem = 97 * 1024 + 1024, iters = 512;
  auto k = rate_kernel<N, A_MN, B_MN, CHAINS, A_TMEM>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k<<<grid, 128, smem>>>(iters, d_out);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); exit(1); }
  long long h[148];
  cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double n_mma = (double)iters * 4 * CHAINS, cyc = mx / n_mma;
  const double bytes = (A_TMEM ? 0 : 4096) + N * 32.0;
  printf("%4d %4d %4d %6d %6d %5d | %7.1f  %8.1f  %8.0f\n", N, A_MN, B_MN, CHAINS, A_TMEM, grid, cyc, bytes / cyc,
         2.0 * 128 * N * 16 / cyc);
}
template <int A_MN, int B_MN, int A_TMEM>
void sweep(int grid) {
  run<32, A_MN, B_MN, 1, A_TMEM>(grid); run<32, A_MN, B_MN, 2, A_TMEM>(grid); run<32, A_MN, B_MN, 4, A_TMEM>(grid);
  run<64, A_MN, B_MN, 1, A_TMEM>(grid); run<64, A_MN, B_MN, 2, A_TMEM>(grid); run<64, A_MN, B_MN, 4, A_TMEM>(grid);
  run<128, A_MN, B_MN, 1, A_TMEM>(grid); run<128, A_MN, B_MN, 2, A_TMEM>(grid);
  run<256, A_MN, B_MN, 1, A_TMEM>(grid);
}

int main() {
  cudaMalloc(&d_out, 1024 * sizeof(long long));
  printf("%4s %4s %4s %6s %6s %5s | cyc/MMA  B/clk(operands)  flop/clk\n", "N", "a_mn", "b_mn", "chains", "a_tmem", "ctas");
  for (int grid : {1, 148}) {
    sweep<0, 0, 0>(grid); sweep<0, 1, 0>(grid); sweep<1, 0, 0>(grid); sweep<1, 1, 0>(grid);
    sweep<0, 0, 1>(grid); sweep<0, 1, 1>(grid);
  }
  return 0;
}
