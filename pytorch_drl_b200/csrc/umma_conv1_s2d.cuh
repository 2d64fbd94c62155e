// conv1 weight gradient on a SPACE-TO-DEPTH view of the frame, with no im2col duplication in shared memory.
//
// An 8x8 stride-4 convolution over [84, 84, 4] is a 2x2 stride-1 convolution over the frame cut into
// 21 x 21 blocks of 4 rows x 4 pixels x 4 channels = 64 values: output (oy, ox) reads blocks
// (oy + ty, ox + tx), ty, tx in {0, 1}.  One CTA handles one sample at a time:
//
//   TMA warp     one bulk copy of the raw uint8 frame (28 224 B, gathered through the minibatch row index)
//                + one tensor-map load of the sample's dY1 [20, 20, 32] bf16 as rows y*21 + x of 64 B
//                (64B swizzle; the x = 20 column is the TMA unit's out-of-bounds zero fill)
//   8 converter  raw frame -> bf16 "block image": row q = by*21 + bx holds the block's 64 values (128 B,
//   warps        128B swizzle), element e = dy*16 + px*4 + ci.  Every pixel is converted ONCE.
//   MMA warp     dW^T[(tap, co)][e] += sum over image rows q of dY[q - off_tap][co] * image[q][e],  off_tap = ty*21 + tx.
//                Both operands are MN-major and read in place: A = the dY tile (64B swizzle), one 32-channel atom per tap, a PAIR
//                of taps per MMA (M = 64; atom 0 = the tap with the larger offset, leading-dimension offset = one dY row; the
//                window of a tap starts `off` rows BEFORE the tile: 22 zero rows lead each tile, 28 follow it), B = the block
//                image (N = 64 elements of a block).  28 k-steps of 16 rows cover q < 448; dY rows with x = 20 or r >= 420 are
//                zeros.  4 KB of operands per MMA; the image-as-A form (M = 128 = two taps x 64 elements, N = 32 channels: 5 KB)
//                it replaces took 0.355 instead of 0.333 ms.
//   epilogue     after the last sample: TMEM (M = 64: rows in lanes 0-15 of each quarter) -> atomicAdd into the OIHW gradient,
//                scaled by 1/255.
//
// Per sample the shared-memory traffic is 28 KB (raw) + 56 KB (image write) + 28 KB raw read + 56 MMAs x 4 KB of operand fetch.
#pragma once
#include <cuda.h>
#include <type_traits>
#include "common.cuh"
#include "umma.cuh"
#include "umma_conv1.cuh"
#include "umma_tma_gemm.cuh"

namespace drl {

constexpr int kS2dRawBytes = 28672;                  // 28 224 B frame, padded to 1 KB
constexpr int kS2dKSteps = 28;                       // 28 x 16 = 448 block-image rows q cover 441 blocks + the largest dY shift
constexpr int kS2dImgRows = 448;                     // 441 block rows, zero tail
constexpr int kS2dImgBytes = kS2dImgRows * 128;      // 57 344 B = 56 KB
// dY region: [22 zero rows][tile 0: 420 rows][28 zero rows][tile 1: 420 rows][28 zero rows] of 64 B.  A tap window reads dY rows
// q - off, off in {0, 1, 21, 22}, q < 448: up to 22 rows before and 28 rows behind a tile, which are zeros (the gap is shared).
constexpr int kS2dDyLead = 22, kS2dDyPitchRows = 448;
constexpr int kS2dDyPitch = kS2dDyPitchRows * 64;    // 28 672 B between the two tiles
constexpr int kS2dDyRegion = (kS2dDyLead + 2 * kS2dDyPitchRows) * 64;   // 58 752 B
constexpr int kS2dWgradSmem = 2 * (kS2dRawBytes + kS2dImgBytes) + kS2dDyRegion + 256 + 1024;
static_assert(kS2dWgradSmem <= 227 * 1024, "conv1 weight-gradient shared-memory plan");
constexpr int kS2dThreads = 32 * (4 + 1 + 8 + 1);    // epilogue, MMA, converters, TMA

struct Conv1S2dArgs {
  const uint8_t* obs;        // [rows, 84, 84, 4]
  const int64_t* row_idx;    // [nb] or nullptr
  int nb;
  float scale;               // float32(1/255)
  float* gw;                 // fp32 [32][4][8][8]
};

__device__ __forceinline__ void tma_load_4d_s2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3,
                                                uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tmap), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(kS2dThreads, 1) conv1_wgrad_s2d_kernel(const __grid_constant__ CUtensorMap tmap_dy,
                                                                          const Conv1S2dArgs args) {
  using namespace umma;
  constexpr int NCONV = 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sImg = smem;                                    // 2 block images
  uint8_t* sDy = smem + 2 * kS2dImgBytes + kS2dDyLead * 64;   // 2 dY tiles (kS2dDyPitch apart) inside the zero-padded dY region
  uint8_t* sRaw = smem + 2 * kS2dImgBytes + kS2dDyRegion;  // 2 raw frames
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + 2 * kS2dRawBytes);
  uint64_t* raw_full = bars;          // [2]
  uint64_t* raw_empty = bars + 2;     // [2]
  uint64_t* img_full = bars + 4;      // [2]
  uint64_t* img_empty = bars + 6;     // [2]  (also frees the dY tile of the same sample)
  uint64_t* dy_full = bars + 8;       // [2]
  uint64_t* done = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // rows the producers never write must read as finite zeros: image rows >= 441 (they only ever meet zero dY
  // rows, but 0 * NaN would poison the sum) and dY rows >= 420
  for (int i = threadIdx.x * 16; i < 2 * kS2dImgBytes + kS2dDyRegion; i += blockDim.x * 16)
    st_shared_v4(smem_u32(smem + i), make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(raw_full + s, 1); mbar_init(raw_empty + s, NCONV / 32);
      mbar_init(img_full + s, NCONV / 32); mbar_init(img_empty + s, 1);
      mbar_init(dy_full + s, 1);
    }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 128);
  if (warp == 13 && lane == 0) tma_prefetch_desc(&tmap_dy);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_mine = (int)blockIdx.x < args.nb ? (args.nb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (n_mine > 0) {
    if (warp == 13) {
      // ================= TMA warp =================
      if (lane == 0) {
        for (int j = 0; j < n_mine; ++j) {
          const int b = blockIdx.x + j * gridDim.x;
          const uint32_t s = j & 1, ph = (j >> 1) & 1;
          mbar_wait(raw_empty + s, ph ^ 1);
          const int64_t srow = args.row_idx ? args.row_idx[b] : (int64_t)b;
          mbar_arrive_expect_tx(raw_full + s, kC1FrameBytes);
          tma_bulk_g2s(sRaw + s * kS2dRawBytes, args.obs + srow * kC1FrameBytes, kC1FrameBytes, raw_full + s);
          mbar_wait(img_empty + s, ph ^ 1);                 // the MMAs of sample j-2 are done with this dY tile
          mbar_arrive_expect_tx(dy_full + s, 420 * 64);
          tma_load_4d_s2d(sDy + s * kS2dDyPitch, &tmap_dy, 0, 0, 0, b, dy_full + s);
        }
      }
      __syncwarp();
    } else if (warp >= 5) {
      // ================= converters: raw uint8 frame -> bf16 block image (every pixel once) ==============
      const int t = threadIdx.x - 160;
      for (int j = 0; j < n_mine; ++j) {
        const uint32_t s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(raw_full + s, ph);
        mbar_wait(img_empty + s, ph ^ 1);
        const uint32_t raw = smem_u32(sRaw + s * kS2dRawBytes), img = smem_u32(sImg + s * kS2dImgBytes);
#pragma unroll
        for (int rep = 0; rep < 2; ++rep) {
          const int q = t + rep * NCONV;                     // block index by*21 + bx
          if (q < 441) {
            const int by = (q * 3121) >> 16;                 // q / 21 for q < 441
            const int bx = q - by * 21;
            const uint32_t src = raw + (uint32_t)(by * 4 * kC1RowBytes + bx * 16);
            uint4 v[4];
#pragma unroll
            for (int dy = 0; dy < 4; ++dy) v[dy] = lds_v4(src + (uint32_t)(dy * kC1RowBytes));
            const uint32_t dst = img + (uint32_t)q * 128u;
#pragma unroll
            for (int dy = 0; dy < 4; ++dy) {
              st_shared_v4(dst + (((uint32_t)(2 * dy) ^ ((uint32_t)q & 7u)) << 4), u8x8_to_bf16x8(make_uint2(v[dy].x, v[dy].y)));
              st_shared_v4(dst + (((uint32_t)(2 * dy + 1) ^ ((uint32_t)q & 7u)) << 4), u8x8_to_bf16x8(make_uint2(v[dy].z, v[dy].w)));
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive_warp(img_full + s);
        mbar_arrive_warp(raw_empty + s);
      }
    } else if (warp == 4) {
      // ================= MMA issuer =================
      // dW^T[(tap, co)][e] += sum_q dY[q - off_tap][co] * image[q][e]: the A operand is the dY tile (MN-major, 64B swizzle; one
      // 32-channel atom per tap), a PAIR of taps per MMA (M = 64: atom 0 = the tap with the larger offset, leading-dimension offset =
      // one dY row), the B operand the block image itself (MN-major, N = 64 elements of a block).  Two accumulators (pairs
      // {(0,1), (0,0)} and {(1,1), (1,0)}) are issued interleaved.  2 KB + 2 KB of operands per MMA instead of 4 KB + 1 KB for the
      // image-as-A form (M = 128 = two taps x 64 elements, N = 32): 20 % less shared-memory operand traffic.
      constexpr uint32_t idesc = make_idesc_f16(1u, 64, 64, 1, 1);      // bf16, MN-major A and B
      constexpr uint32_t a_hi = desc_hi(512, LAYOUT_SW64), b_hi = desc_hi(1024, LAYOUT_SW128);
      const uint32_t a_base0 = desc_lo(smem_u32(sDy) - 1 * 64, 64), a_base1 = desc_lo(smem_u32(sDy) - 22 * 64, 64);
      const uint32_t b_base = desc_lo(smem_u32(sImg), 0);
      auto issue = [&](uint32_t a_off, uint32_t b_off, auto first) {
        constexpr bool FIRST = decltype(first)::value;
#pragma unroll
        for (int k = 0; k < kS2dKSteps; ++k) {
          const uint32_t bl = b_base + b_off + k * ((16 * 128) >> 4);
          const uint32_t ak = a_off + k * ((16 * 64) >> 4);
          if (FIRST && k == 0) {
            mma_f16_lohi<false>(tmem_base, a_base0 + ak, a_hi, bl, b_hi, idesc);
            mma_f16_lohi<false>(tmem_base + 64, a_base1 + ak, a_hi, bl, b_hi, idesc);
          } else {
            mma_f16_lohi<true>(tmem_base, a_base0 + ak, a_hi, bl, b_hi, idesc);
            mma_f16_lohi<true>(tmem_base + 64, a_base1 + ak, a_hi, bl, b_hi, idesc);
          }
        }
      };
      for (int j = 0; j < n_mine; ++j) {
        const uint32_t s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(img_full + s, ph);
        mbar_wait(dy_full + s, ph);
        tc_fence_after();
        if (elect_one()) {
          if (j == 0) issue(s * (kS2dDyPitch >> 4), s * (kS2dImgBytes >> 4), std::true_type{});
          else issue(s * (kS2dDyPitch >> 4), s * (kS2dImgBytes >> 4), std::false_type{});
          mma_commit(img_empty + s);
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(done);
      __syncwarp();
    } else {
      // ================= epilogue (once): TMEM -> OIHW gradient =================
      // M = 64 accumulators keep their rows in lanes 0-15 of every 32-lane quarter: row m = 16 * quarter + lane = (tap of the pair, co)
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + pr * 64 + half * 32, r);
          tmem_ld_wait();
          if (lane < 16) {
            const int m = 16 * warp + lane, tp = m >> 5, co = m & 31;
            const int ty = pr, tx = 1 - tp;                  // pair pr: atom 0 = tap (ty, 1), atom 1 = tap (ty, 0)
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int e = half * 32 + c;
              const int ky = 4 * ty + (e >> 4), kx = 4 * tx + ((e >> 2) & 3), ci = e & 3;
              atomicAdd(args.gw + ((co * 4 + ci) * 8 + ky) * 8 + kx, __uint_as_float(r[c]) * args.scale);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace drl
