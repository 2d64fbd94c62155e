// Shared pieces of the conv1 kernels (8x8 stride 4 over uint8 84x84x4 frames): frame geometry, shared-memory loads, the exact
// uint8 -> bf16 conversion of the weight-gradient kernels.  The kernels are in umma_conv1_i8.cuh (NatureCNN forward: integer MMAs
// on the raw bytes), umma_conv1_s2d.cuh (NatureCNN weight gradient on a space-to-depth block image) and umma_rnd_conv1.cuh
// (RND predictor); the x*(1/255) of scale_observations.py:35 is applied to the accumulator.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace drl {

constexpr int kC1FrameBytes = 84 * 84 * 4;
constexpr int kC1RowBytes = 84 * 4;

__device__ __forceinline__ uint2 lds_v2(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}

__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// 8 uint8 -> 8 bf16 (exact: 0..255 fit the 8-bit significand): fp32 magic number then truncation.
__device__ __forceinline__ uint4 u8x8_to_bf16x8(uint2 w) {
  const uint32_t magic = 0x4B000000u;                 // 2^23
  float f[8];
  f[0] = __uint_as_float(__byte_perm(w.x, magic, 0x7440)) - 8388608.f;   // {b0,0,0,0x4B}
  f[1] = __uint_as_float(__byte_perm(w.x, magic, 0x7441)) - 8388608.f;
  f[2] = __uint_as_float(__byte_perm(w.x, magic, 0x7442)) - 8388608.f;
  f[3] = __uint_as_float(__byte_perm(w.x, magic, 0x7443)) - 8388608.f;
  f[4] = __uint_as_float(__byte_perm(w.y, magic, 0x7440)) - 8388608.f;
  f[5] = __uint_as_float(__byte_perm(w.y, magic, 0x7441)) - 8388608.f;
  f[6] = __uint_as_float(__byte_perm(w.y, magic, 0x7442)) - 8388608.f;
  f[7] = __uint_as_float(__byte_perm(w.y, magic, 0x7443)) - 8388608.f;
  uint4 o;   // bf16x2 = high halves of the two floats
  o.x = __byte_perm(__float_as_uint(f[0]), __float_as_uint(f[1]), 0x7632);
  o.y = __byte_perm(__float_as_uint(f[2]), __float_as_uint(f[3]), 0x7632);
  o.z = __byte_perm(__float_as_uint(f[4]), __float_as_uint(f[5]), 0x7632);
  o.w = __byte_perm(__float_as_uint(f[6]), __float_as_uint(f[7]), 0x7632);
  return o;
}

}  // namespace drl
