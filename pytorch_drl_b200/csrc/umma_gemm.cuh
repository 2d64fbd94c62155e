// Shared enums of the tcgen05 kernels of the bf16 engine (umma_tma_*.cuh, umma_rnd*.cuh).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace drl {

// Epilogue kinds.  "MASK" epilogues are data gradients: the accumulator is multiplied by the derivative of the
// activation of the layer below, read as 1 bit per element (umma.cuh relu_bit_pos) or derived from the stored activation.
enum {
  EPI_BIAS_RELU_BF16 = 0,    // y = relu(acc * scale + bias)        -> bf16 (+ mask bits)
  EPI_BIAS_RELU_F32 = 1,     // y = relu(acc + bias)                -> fp32
  EPI_MASK_BF16 = 2,         // g = bit ? acc : 0                   -> bf16 (+ fused bias gradient of the layer below)
  EPI_BIAS_F32 = 3,          // y = acc + bias                      -> fp32            (RND: last linear layers)
  EPI_BIAS_LEAKY_BF16 = 4,   // y = leaky_relu(acc + bias, 0.2)     -> bf16 (+ sign bits)   (RND trunk)
  EPI_MASK_LEAKY_BF16 = 5    // g = bit ? acc : 0.2 * acc           -> bf16 (+ fused bias gradient)
};
__host__ __device__ constexpr bool epi_is_mask(int epi) { return epi == EPI_MASK_BF16 || epi == EPI_MASK_LEAKY_BF16; }
__host__ __device__ constexpr bool epi_is_leaky(int epi) { return epi == EPI_BIAS_LEAKY_BF16 || epi == EPI_MASK_LEAKY_BF16; }
constexpr float kLeakySlope = 0.2f;      // random_network_distillation.py:31-35

}  // namespace drl
