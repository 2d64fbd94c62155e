// Shared pieces of the conv1 kernels of NatureCNN (8x8 stride 4 over uint8 84x84x4 frames -> 20x20x32): argument block,
// exact uint8 -> fp16 / bf16 conversions, the row epilogue.  The kernels themselves are in umma_conv1_s2d.cuh (space-to-depth
// block image); the x*(1/255) of scale_observations.py:35 is applied to the accumulator.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace drl {

struct Conv1Args {
  const uint8_t* obs;        // [rows, 84, 84, 4]
  const int64_t* row_idx;    // [nb] or nullptr
  int nb;
  const void* bpack;         // fwd: fp16 [4][32 x 64] swizzled tile images (16 KB)
  const float* bias;         // fwd: [32] EFFECTIVE bias = b - 1024 * scale * sum_k fp16(W[n][k]) (operands hold 1024 + pixel)
  float scale;               // float32(1/255)
  __nv_bfloat16* out;        // fwd: act1 [nb*400, 32]
  uint32_t* mask_bits;       // fwd: ReLU mask word of position m (umma.cuh relu_bit_pos): what conv2's data gradient needs
  const __nv_bfloat16* dy;   // wgrad: dY1 [nb*400, 32]
  float* gw;                 // wgrad: fp32 [32][4][8][8]
};

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

// 8 uint8 -> 8 fp16 holding 1024 + b (bytes become the mantissa of 0x6400 = 1024.0): one PRMT per pair.
// The constant 1024 * sum_k W[n][k] this adds to every accumulator is folded into the bias
// (last block of pack_all_kernel), so the forward conversion costs 4 instructions per 8 pixels.
__device__ __forceinline__ uint4 u8x8_to_f16x8_biased(uint2 w) {
  const uint32_t magic = 0x64006400u;
  return make_uint4(__byte_perm(w.x, magic, 0x7170), __byte_perm(w.x, magic, 0x7372),
                    __byte_perm(w.y, magic, 0x7170), __byte_perm(w.y, magic, 0x7372));
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

// conv1 epilogue of one output position: relu(acc * (1/255) + bias) -> 32 bf16 (one 64-byte row) + its ReLU mask word
__device__ __forceinline__ void c1_store_row(const uint32_t (&r)[32], const float (&bias)[32], float scale,
                                             __nv_bfloat16* out_row, uint32_t* bits_dst) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    pk[i] = umma::pack_bf16x2_relu(fmaf(__uint_as_float(r[2 * i]), scale, bias[2 * i]),
                                   fmaf(__uint_as_float(r[2 * i + 1]), scale, bias[2 * i + 1]));
#pragma unroll
  for (int q = 0; q < 2; ++q)
    umma::stg_v8(out_row + 16 * q, pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3], pk[8 * q + 4], pk[8 * q + 5],
                 pk[8 * q + 6], pk[8 * q + 7]);
  *bits_dst = umma::relu_bits_from_pairs(pk);
}


}  // namespace drl
