// First convolution of the RND networks (random_network_distillation.py:27-30: Conv2d(1, 16, 8x8, stride 4) + LeakyReLU(0.2)
// on the LAST channel of the frame, normalised and clipped, :228-229; normalize.py:149-162) on tcgen05.
//
// The input is a real-valued image (normaliser output), so the operand has to be produced by converter warps anyway:
//   TMA warp      bulk copies of the raw uint8 frames of a SAMPLE PAIR (gathered through the minibatch row index)
//   8 converter   phase A: last channel -> z = clamp(x * a[px] + b[px], -5, 5) -> bf16 image [84][84] per frame (every pixel
//   warps                  once; a = rsqrt(m2 - m1^2 + eps) / 255, b = -m1 * rsqrt(..) come from rnd_norm_table_kernel)
//                 phase B: im2col rows: GEMM row r = oy*20 + ox holds its 8x8 patch as 8 chunks of 16 B (chunk = ky), written
//                          into 128B-swizzled [128 x 64] tiles — one 128-byte row per output position
//   MMA warp      forward:  D[128 x 32] = tile x [teacher W1 (16) | student W1 (16)]^T for both frames of the pair (two chains)
//                 wgrad:    dW1^T[(frame, ky, kx)][(s, co)] += tile^T x dY1  — the tiles of the two frames are the two 64-row atoms of
//                           one MN-major A operand; dY1 is the pair-packed [400][2 x 16] gradient (TMA, 64B swizzle)
//   epilogue      forward:  bias + LeakyReLU, PAIR-PACKED outputs: teacher act1 [pairs][400][2][16] and student act1 likewise
//                           (+ sign bits), so that the next layers run both samples of a pair through ONE block-diagonal GEMM
//                 wgrad:    after the last pair: the diagonal blocks (frame == s) are added to the OIHW gradient
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"
#include "umma_conv1.cuh"
#include "umma_tma_gemm.cuh"

namespace drl {

struct RndConv1Args {
  const uint8_t* obs;        // [rows, 84, 84, 4] uint8
  const int64_t* row_idx;    // [nb] or nullptr
  int nb;
  const float* na;           // [7056] per-pixel scale
  const float* nbias;        // [7056] per-pixel shift
  const void* bpack;         // fwd: bf16 [32 x 64] swizzled tile image: rows 0-15 teacher, 16-31 student; k = ky*8 + kx
  const float* bias;         // fwd: [32] = (teacher b1 | student b1)
  __nv_bfloat16* tout;       // fwd: teacher act1, pair-packed [pairs][400][32] (channel = s*16 + c)
  __nv_bfloat16* sout;       // fwd: student act1, same layout
  uint32_t* sbits;           // fwd: sign bits of sout, one word per row (nullable)
  float* gw;                 // wgrad: student dW1 fp32 [16][64]
};

constexpr int kR1Steps = 2;                          // ring of tile steps (one step = tile t of both frames)
constexpr int kR1RawStages = 2;                      // raw frame pairs in flight (the kernels are bound by this HBM read)
constexpr int kR1TileBytes = 128 * 128;
constexpr int kR1StepBytes = 2 * kR1TileBytes;
constexpr int kR1ImgBytes = 14336;                   // 84 rows x 168 B = 14 112, padded
constexpr int kR1RawBytes = 28672;                   // 28 224 B frame, padded to 1 KB
constexpr int kR1DyBytes = 128 * 64;                 // wgrad: dY tile of a step (128 rows x 32 bf16)
constexpr int kR1NAcc = 4;                           // fwd accumulator ring (64 columns each)
constexpr int kR1Threads = 32 * (4 + 1 + 8 + 1);     // wgrad: epilogue, MMA, converters, TMA
constexpr int kR1FEpiWG = 2;                         // forward: epilogue warpgroups (steps round-robin)
constexpr int kR1FThreads = 32 * (4 * kR1FEpiWG + 1 + 8 + 1);
constexpr int kR1FwdSmem = 4096 + kR1Steps * kR1StepBytes + 2 * kR1ImgBytes + kR1RawStages * 2 * kR1RawBytes + 256 + 1024;
constexpr int kR1WgradSmem = kR1Steps * (kR1StepBytes + kR1DyBytes) + 2 * kR1ImgBytes + kR1RawStages * 2 * kR1RawBytes + 256 + 1024;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tmap), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// per-pixel affine form of (x * (1/255) - m1) * rsqrt(m2 - m1^2 + eps)   (normalize.py:149-152)
__global__ void __launch_bounds__(256) rnd_norm_table_kernel(const float* __restrict__ m1, const float* __restrict__ m2,
                                                             float* __restrict__ na, float* __restrict__ nbias, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mu = m1[i];
  const float rs = rsqrtf(m2[i] - mu * mu + 1e-8f);
  na[i] = rs * __uint_as_float(0x3b808081u);         // float32(1/255), scale_observations.py:35
  nbias[i] = -mu * rs;
}

// ---- converter warps (shared by forward and weight gradient) ---------------------------------------------------------
struct R1Smem {
  uint8_t* sA;      // kR1Steps x 2 tiles
  uint8_t* sImg;    // 2 images
  uint8_t* sRaw;    // kR1RawStages x 2 raw frames
  uint64_t* raw_full; uint64_t* raw_empty; uint64_t* a_full; uint64_t* a_empty;
};

__device__ __forceinline__ void r1_converters(const RndConv1Args& args, const R1Smem& sm, int n_pairs_mine, int t /* 0..255 */) {
  using namespace umma;
  const uint32_t img = smem_u32(sm.sImg), raw0 = smem_u32(sm.sRaw);
  // Thread t owns the pixel quads q = t + 256 i of EVERY frame, so its slice of the per-pixel scale / shift tables lives in
  // registers for the whole kernel (a global load per task inside the loop made the converters latency bound: 0.80 ms)
  float4 ta[7], tb[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) {
    const int q = t + 256 * i;
    ta[i] = q < 1764 ? __ldg(reinterpret_cast<const float4*>(args.na) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    tb[i] = q < 1764 ? __ldg(reinterpret_cast<const float4*>(args.nbias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int j = 0; j < n_pairs_mine; ++j) {
    const int p = blockIdx.x + j * gridDim.x;
    const bool second = 2 * p + 1 < args.nb;                 // odd batch: the last pair has one real frame
    const uint32_t rs = (uint32_t)j % kR1RawStages, raw = raw0 + rs * 2 * kR1RawBytes;
    mbar_wait(sm.raw_full + rs, (j / kR1RawStages) & 1);
    // phase A: normalise 4 pixels per task (one 16-byte word of the NHWC frame: 4 pixels x 4 channels)
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      const int q = t + 256 * i;
      if (q < 1764) {
        uint4 w[2];
        w[0] = lds_v4(raw + (uint32_t)(q * 16));
        w[1] = second ? lds_v4(raw + (uint32_t)(kR1RawBytes + q * 16)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          // byte 3 of each word -> float exactly (PRMT into the mantissa of 2^23, minus 2^23), affine, bf16 pair, clamp to
          // [-5, 5] in bf16x2 (both bounds are exact in bf16, so clamping after the rounding gives the same value)
          const float x0 = __uint_as_float(__byte_perm(w[f].x, 0x4B000000u, 0x7443)) - 8388608.f;
          const float x1 = __uint_as_float(__byte_perm(w[f].y, 0x4B000000u, 0x7443)) - 8388608.f;
          const float x2 = __uint_as_float(__byte_perm(w[f].z, 0x4B000000u, 0x7443)) - 8388608.f;
          const float x3 = __uint_as_float(__byte_perm(w[f].w, 0x4B000000u, 0x7443)) - 8388608.f;
          __nv_bfloat162 p0 = __floats2bfloat162_rn(fmaf(x0, ta[i].x, tb[i].x), fmaf(x1, ta[i].y, tb[i].y));
          __nv_bfloat162 p1 = __floats2bfloat162_rn(fmaf(x2, ta[i].z, tb[i].z), fmaf(x3, ta[i].w, tb[i].w));
          const __nv_bfloat162 lo = __floats2bfloat162_rn(-5.f, -5.f), hi = __floats2bfloat162_rn(5.f, 5.f);
          p0 = __hmin2(__hmax2(p0, lo), hi);
          p1 = __hmin2(__hmax2(p1, lo), hi);
          uint2 o = make_uint2(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1));
          if (f == 1 && !second) o = make_uint2(0u, 0u);     // phantom frame: zeros
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(img + (uint32_t)(f * kR1ImgBytes + q * 8)), "r"(o.x), "r"(o.y) : "memory");
        }
      }
    }
    mbar_arrive_warp(sm.raw_empty + rs);
    asm volatile("bar.sync 1, 256;" ::: "memory");           // both images complete
    // phase B: im2col tiles, one step = tile `tile` of both frames
#pragma unroll 1
    for (int tile = 0; tile < 4; ++tile) {
      const uint32_t u = (uint32_t)(4 * j + tile), slot = u % kR1Steps;
      mbar_wait(sm.a_empty + slot, ((u / kR1Steps) & 1) ^ 1);
      // thread = (frame f = t >> 7, tile row = t & 127): the 8 chunks (ky) of its im2col row are 8 image rows 168 B apart
      // (immediate offsets) and land at rowbase ^ (ky << 4) of the 128B-swizzled tile: 2 LDS.64 + 1 LOP + 1 STS.128 per chunk
      const int row = t & 127, f = t >> 7;
      const int pos = tile * 128 + row;
      if (pos < 400) {
        const int oy = (pos * 3277) >> 16, ox = pos - oy * 20;
        const uint32_t src = img + (uint32_t)(f * kR1ImgBytes + oy * (4 * 168) + ox * 8);
        const uint32_t dst = smem_u32(sm.sA) + slot * kR1StepBytes + (uint32_t)(f * kR1TileBytes + row * 128 + ((row & 7) << 4));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint2 lo = lds_v2(src + c * 168), hi = lds_v2(src + c * 168 + 8);
          st_shared_v4(dst ^ (uint32_t)(c << 4), make_uint4(lo.x, lo.y, hi.x, hi.y));
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(sm.a_full + slot);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");           // everybody is done reading the images
  }
}

// TMA of the raw frames of the CTA's pairs
__device__ __forceinline__ void r1_load_raw(const RndConv1Args& args, const R1Smem& sm, int j) {
  using namespace umma;
  const int p = blockIdx.x + j * gridDim.x;
  const uint32_t rs = (uint32_t)j % kR1RawStages;
  mbar_wait(sm.raw_empty + rs, ((j / kR1RawStages) & 1) ^ 1);
  const bool second = 2 * p + 1 < args.nb;
  mbar_arrive_expect_tx(sm.raw_full + rs, (uint32_t)(kC1FrameBytes * (second ? 2 : 1)));
  for (int f = 0; f < (second ? 2 : 1); ++f) {
    const int b = 2 * p + f;
    const int64_t srow = args.row_idx ? args.row_idx[b] : (int64_t)b;
    tma_bulk_g2s(sm.sRaw + (rs * 2 + f) * kR1RawBytes, args.obs + srow * kC1FrameBytes, kC1FrameBytes, sm.raw_full + rs);
  }
}

// ---- forward -----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kR1FThreads, 1) rnd_conv1_fwd_kernel(const RndConv1Args args) {
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  R1Smem sm;
  sm.sA = smem + 4096;
  sm.sImg = sm.sA + kR1Steps * kR1StepBytes;
  sm.sRaw = sm.sImg + 2 * kR1ImgBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm.sRaw + kR1RawStages * 2 * kR1RawBytes);
  sm.raw_full = bars; sm.raw_empty = bars + 2; sm.a_full = bars + 4; sm.a_empty = bars + 4 + kR1Steps;
  uint64_t* acc_full = bars + 4 + 2 * kR1Steps;      // [kR1NAcc]
  uint64_t* acc_empty = acc_full + kR1NAcc;          // [kR1NAcc]
  uint64_t* b_full = acc_empty + kR1NAcc;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  constexpr int W_MMA = 4 * kR1FEpiWG, W_CONV = W_MMA + 1, W_TMA = W_CONV + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = (args.nb + 1) / 2;
  const int n_mine = (int)blockIdx.x < n_pairs ? (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int q = 0; q < kR1RawStages; ++q) { mbar_init(sm.raw_full + q, 1); mbar_init(sm.raw_empty + q, 8); }
    for (int s = 0; s < kR1Steps; ++s) { mbar_init(sm.a_full + s, 8); mbar_init(sm.a_empty + s, 1); }
    for (int a = 0; a < kR1NAcc; ++a) { mbar_init(acc_full + a, 1); mbar_init(acc_empty + a, 4); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0 && n_mine > 0) {
      mbar_arrive_expect_tx(b_full, 4096);
      tma_bulk_g2s(sB, args.bpack, 4096, b_full);
      for (int j = 0; j < n_mine; ++j) r1_load_raw(args, sm, j);
    }
    __syncwarp();
  } else if (warp >= W_CONV) {
    r1_converters(args, sm, n_mine, threadIdx.x - 32 * W_CONV);
  } else if (warp == W_MMA) {
    constexpr uint32_t idesc = make_idesc_f16(1u, 128, 32, 0, 0);
    constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
    const uint32_t b_lo = desc_lo(smem_u32(sB), 0), a_lo0 = desc_lo(smem_u32(sm.sA), 0);
    if (n_mine > 0) mbar_wait(b_full, 0);
    for (uint32_t u = 0; u < (uint32_t)(4 * n_mine); ++u) {
      const uint32_t slot = u % kR1Steps, acc = u % kR1NAcc;
      mbar_wait(sm.a_full + slot, (u / kR1Steps) & 1);
      mbar_wait(acc_empty + acc, ((u / kR1NAcc) & 1) ^ 1);
      tc_fence_after();
      const uint32_t a0 = a_lo0 + slot * (kR1StepBytes >> 4), a1 = a0 + (kR1TileBytes >> 4);
      const uint32_t d0 = tmem_base + acc * 64, d1 = d0 + 32;
      if (elect_one()) {
        mma_f16_lohi<false>(d0, a0, hi, b_lo, hi, idesc);
        mma_f16_lohi<false>(d1, a1, hi, b_lo, hi, idesc);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
          mma_f16_lohi<true>(d0, a0 + 2 * k, hi, b_lo + 2 * k, hi, idesc);
          mma_f16_lohi<true>(d1, a1 + 2 * k, hi, b_lo + 2 * k, hi, idesc);
        }
        mma_commit(acc_full + acc);
        mma_commit(sm.a_empty + slot);
      }
      __syncwarp();
    }
  } else {
    float bias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias[i] = __ldg(args.bias + i);
    const int row = threadIdx.x & 127, wg = warp >> 2;
    for (uint32_t u = 0; u < (uint32_t)(4 * n_mine); ++u) {
      if ((int)(u % kR1FEpiWG) != wg) continue;
      const uint32_t acc = u % kR1NAcc;
      const int j = (int)(u >> 2), tile = (int)(u & 3);
      const int p = blockIdx.x + j * gridDim.x;
      const int pos = tile * 128 + row;
      mbar_wait(acc_full + acc, (u / kR1NAcc) & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * 64;
      tmem_ld_32x32b_x32(ta, r0);
      tmem_ld_32x32b_x32(ta + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_warp(acc_empty + acc);
      if (pos < 400) {
        auto act = [&](uint32_t raw_acc, int i) {
          const float v = __uint_as_float(raw_acc) + bias[i];
          return fmaxf(v, kLeakySlope * v);
        };
        uint32_t pt[16], ps[16];                      // teacher row [f0: 16 | f1: 16], student row likewise
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pt[i] = pack_bf16x2(act(r0[2 * i], 2 * i), act(r0[2 * i + 1], 2 * i + 1));
          pt[8 + i] = pack_bf16x2(act(r1[2 * i], 2 * i), act(r1[2 * i + 1], 2 * i + 1));
          ps[i] = pack_bf16x2(act(r0[16 + 2 * i], 16 + 2 * i), act(r0[16 + 2 * i + 1], 16 + 2 * i + 1));
          ps[8 + i] = pack_bf16x2(act(r1[16 + 2 * i], 16 + 2 * i), act(r1[16 + 2 * i + 1], 16 + 2 * i + 1));
        }
        const int64_t m = (int64_t)p * 400 + pos;
        __nv_bfloat16* to = args.tout + m * 32;
        __nv_bfloat16* so = args.sout + m * 32;
        stg_v8(to, pt[0], pt[1], pt[2], pt[3], pt[4], pt[5], pt[6], pt[7]);
        stg_v8(to + 16, pt[8], pt[9], pt[10], pt[11], pt[12], pt[13], pt[14], pt[15]);
        stg_v8(so, ps[0], ps[1], ps[2], ps[3], ps[4], ps[5], ps[6], ps[7]);
        stg_v8(so + 16, ps[8], ps[9], ps[10], ps[11], ps[12], ps[13], ps[14], ps[15]);
        if (args.sbits) args.sbits[m] = pos_bits_from_pairs(ps);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// ---- weight gradient -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kR1Threads, 1) rnd_conv1_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_dy,
                                                                         const RndConv1Args args) {
  using namespace umma;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  R1Smem sm;
  sm.sA = smem;
  uint8_t* sDy = smem + kR1Steps * kR1StepBytes;
  sm.sImg = sDy + kR1Steps * kR1DyBytes;
  sm.sRaw = sm.sImg + 2 * kR1ImgBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm.sRaw + kR1RawStages * 2 * kR1RawBytes);
  sm.raw_full = bars; sm.raw_empty = bars + 2; sm.a_full = bars + 4; sm.a_empty = bars + 4 + kR1Steps;
  uint64_t* dy_full = bars + 4 + 2 * kR1Steps;       // [kR1Steps]
  uint64_t* done = dy_full + kR1Steps;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_pairs = (args.nb + 1) / 2;
  const int n_mine = (int)blockIdx.x < n_pairs ? (n_pairs - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    for (int q = 0; q < kR1RawStages; ++q) { mbar_init(sm.raw_full + q, 1); mbar_init(sm.raw_empty + q, 8); }
    for (int s = 0; s < kR1Steps; ++s) { mbar_init(sm.a_full + s, 8); mbar_init(sm.a_empty + s, 1); mbar_init(dy_full + s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 32);
  if (warp == 13 && lane == 0) tma_prefetch_desc(&tmap_dy);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (n_mine > 0) {
    if (warp == 13) {
      if (lane == 0) {
        // raw frames of pair j, then the four dY tiles of pair j (each waits for its ring slot)
        r1_load_raw(args, sm, 0);
        for (int j = 0; j < n_mine; ++j) {
          if (j + 1 < n_mine) r1_load_raw(args, sm, j + 1);      // the next pair's frames are in flight while this pair is processed
          const int p = blockIdx.x + j * gridDim.x;
          for (int tile = 0; tile < 4; ++tile) {
            const uint32_t u = (uint32_t)(4 * j + tile), slot = u % kR1Steps;
            mbar_wait(sm.a_empty + slot, ((u / kR1Steps) & 1) ^ 1);
            mbar_arrive_expect_tx(dy_full + slot, kR1DyBytes);
            tma_load_3d(sDy + slot * kR1DyBytes, &tmap_dy, 0, tile * 128, p, dy_full + slot);   // rows >= 400: zero fill
          }
        }
      }
      __syncwarp();
    } else if (warp >= 5) {
      r1_converters(args, sm, n_mine, threadIdx.x - 160);
    } else if (warp == 4) {
      constexpr uint32_t idesc = make_idesc_f16(1u, 128, 32, 1, 1);       // bf16, MN-major A and B
      constexpr uint32_t a_hi = desc_hi(1024, LAYOUT_SW128), b_hi = desc_hi(512, LAYOUT_SW64);
      const uint32_t a_lo0 = desc_lo(smem_u32(sm.sA), kR1TileBytes);       // second 64-row atom = the other frame's tile
      const uint32_t b_lo0 = desc_lo(smem_u32(sDy), 0);
      for (uint32_t u = 0; u < (uint32_t)(4 * n_mine); ++u) {
        const uint32_t slot = u % kR1Steps;
        mbar_wait(sm.a_full + slot, (u / kR1Steps) & 1);
        mbar_wait(dy_full + slot, (u / kR1Steps) & 1);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + slot * (kR1StepBytes >> 4), b_lo = b_lo0 + slot * (kR1DyBytes >> 4);
        const bool last_tile = (u & 3) == 3;                               // positions 384..399 only: one 16-row step
        if (elect_one()) {
          if (u == 0) mma_f16_lohi<false>(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc);
          else mma_f16_lohi<true>(tmem_base, a_lo, a_hi, b_lo, b_hi, idesc);
          if (!last_tile) {
#pragma unroll
            for (int k = 1; k < 8; ++k)
              mma_f16_lohi<true>(tmem_base, a_lo + k * ((16 * 128) >> 4), a_hi, b_lo + k * ((16 * 64) >> 4), b_hi, idesc);
          }
          mma_commit(sm.a_empty + slot);
        }
        __syncwarp();
      }
      if (elect_one()) mma_commit(done);
      __syncwarp();
    } else {
      mbar_wait(done, 0);
      tc_fence_after();
      const int i = threadIdx.x, f = i >> 6, kp = i & 63;       // accumulator row = (frame, ky*8 + kx)
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16), r);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float v = f == 0 ? __uint_as_float(r[c]) : __uint_as_float(r[16 + c]);   // diagonal block: frame == s
        atomicAdd(args.gw + c * 64 + kp, v);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

}  // namespace drl
