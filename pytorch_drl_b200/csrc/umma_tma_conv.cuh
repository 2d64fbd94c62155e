// Family 1c — `tma_conv_kernel`: implicit-GEMM convolution with NO im2col materialisation.
//
// One tile = one sample.  Its NHWC bf16 activation (zero-padded by the TMA unit's out-of-bounds
// fill where the convolution needs padding) lands in shared memory with ONE 4-D tensor-map TMA load
// as a stack of "pixel rows" of 128 bytes (64 channels, or a pair of 32-channel pixels), 128B-swizzled.
// The GEMM rows are indexed on the INPUT-width grid (row r = y * RW + x with RW the staged row width),
// so for tap (ty, tx) the A operand is simply the staged image shifted by ty*RW + tx pixel rows:
// the MMA reads it in place through a shared-memory descriptor whose start address is advanced by
// the shift.  (Measured on B200: the 128B swizzle XOR uses the ABSOLUTE smem address bits [7,10), so a
// start address that is not 1024-byte aligned needs no matrix-base-offset: scripts/try_base_offset.py.)  Rows whose x falls
// outside the output width are computed and discarded (51-82 % of the rows are useful), which is
// cheap because these small-N layers are operand-bandwidth bound, not tensor bound.
// All weight k-block tiles of the layer stay resident in shared memory (one bulk copy per CTA).
// No producer warps, no generic-proxy writes, no proxy fences: warps 0-3 epilogue, 4 MMA, 5 TMA.
#pragma once
#include <cuda.h>
#include <type_traits>
#include "umma_tma_gemm.cuh"

namespace drl {

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tmap), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct TmaConvArgs {
  int nb;
  int KB, KBX;            // taps (k-blocks of 64) and taps per tap-row
  int RW;                 // GEMM row pitch: row r = y * RW + x
  int SHIFT_RW;           // staged pixel rows per tap row: tap (ty, tx) shifts the image by ty*SHIFT_RW + tx
  int pad;                // TMA start coordinate = (-pad, -pad)
  int planes;             // 1, or 2: stride-2 conv, even / odd input rows staged as separate planes (tap row ty
                          //    reads plane ty & 1 shifted by (ty >> 1) * SHIFT_RW + tx)
  int plane_bytes;        // smem pitch of one plane (box bytes rounded up to 1024)
  int box_bytes;          // bytes one sample's TMA box writes
  int raw_buf_bytes;      // staging buffer pitch (planes * plane_bytes, multiple of 1024)
  int slack_bytes;        // readable tail after the last staging buffer (window over-read of discarded rows)
  int MT;                 // M tiles (of 128 GEMM rows) per sample
  int OH, OW;             // valid output grid
  const void* bpack;      // [KB][N x 64] swizzled weight tile images
  // epilogue (same conventions as RowGatherArgs)
  void* out; const float* bias; float scale; const __nv_bfloat16* mask_src;
  int64_t out_sample_pitch; int out_row_pitch; int out_C; int osy, osx; int classes;
  int use_base_offset;    // debug switch for the descriptor base-offset field
  uint32_t* bits_out;     // EPI_BIAS_RELU: ReLU mask words (umma.cuh relu_bit_pos), word e >> 5 for output element index e
  const uint32_t* bits_in; // EPI_MASK: the same bit layout for the tensor shaped like `out` (replaces mask_src reads)
  float* bias_grad;       // EPI_MASK: += column sums of the produced gradient, index = column % bias_mod (or nullptr);
  int bias_mod;           //    bias_mod may be any divisor of the kernel's 32 / 64-column bias group (pair-packed samples fold)
  int chunk_pitch;        // elements between the 32-column chunks of an output row (0 = 32: contiguous).  RND conv3 forward
                          //    writes the two samples of a pair (chunks 0 / 1) to separate per-sample [36][32] blocks.
};

// One accumulator tile of the row epilogue: GEMM row r of sample b -> bias + ReLU (+ mask bits) or ReLU-mask of the
// layer below (+ running bias-gradient column sums).  `tmem_acc` = this warp's lane quarter, column 0 of the
// accumulator; `release()` frees the accumulator once its last chunk sits in registers.  Shared by the single-CTA
// kernel and the CTA-pair kernel (umma_tma_conv2.cuh).
template <int N, int EPI, int BG, int CS, typename Release>
__device__ __forceinline__ void tma_conv_epilogue_tile(const TmaConvArgs& args, int b, int r, bool sample_ok, uint32_t tmem_acc,
                                                       const float* sBias, float (&cs)[CS], uint64_t* tfull_bar,
                                                       uint32_t tfull_parity, Release release) {
  using namespace umma;
  auto coff = [&](int ch) -> int64_t {
    return args.classes ? (int64_t)(ch >> 1) * args.out_row_pitch + (int64_t)(ch & 1) * args.out_C
                        : (int64_t)ch * (args.chunk_pitch ? args.chunk_pitch : 32);
  };
  const int y = r / args.RW, x = r - y * args.RW;
  const bool ok = sample_ok && y < args.OH && x < args.OW;
  const int64_t ooff = (int64_t)b * args.out_sample_pitch + (int64_t)(y * args.osy) * args.out_row_pitch +
                       (int64_t)(x * args.osx) * args.out_C;
  uint32_t mbits[N / 32];                           // ReLU mask bits of this row, one word per 32-column chunk
  if (epi_is_mask(EPI) && ok) {
#pragma unroll
    for (int ch = 0; ch < N / 32; ++ch) mbits[ch] = __ldg(args.bits_in + ((ooff + coff(ch)) >> 5));
  }
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after();
#pragma unroll
  for (int ch = 0; ch < N / 32; ++ch) {
    uint32_t rr[32];
    tmem_ld_32x32b_x32(tmem_acc + ch * 32, rr);
    tmem_ld_wait();
    if (ch == N / 32 - 1) {
      tc_fence_before();
      release();
    }
    if (ok) {
      if (EPI == EPI_BIAS_RELU_BF16 || EPI == EPI_BIAS_LEAKY_BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + ooff + coff(ch);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float v0 = fmaf(__uint_as_float(rr[2 * i]), args.scale, sBias[ch * 32 + 2 * i]);
          const float v1 = fmaf(__uint_as_float(rr[2 * i + 1]), args.scale, sBias[ch * 32 + 2 * i + 1]);
          pk[i] = EPI == EPI_BIAS_RELU_BF16 ? pack_bf16x2_relu(v0, v1)
                                            : pack_bf16x2(fmaxf(v0, kLeakySlope * v0), fmaxf(v1, kLeakySlope * v1));
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
          stg_v8(o + 16 * q, pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3], pk[8 * q + 4], pk[8 * q + 5], pk[8 * q + 6],
                 pk[8 * q + 7]);
        if (args.bits_out) args.bits_out[(ooff + coff(ch)) >> 5] = EPI == EPI_BIAS_RELU_BF16 ? relu_bits_from_pairs(pk) : pos_bits_from_pairs(pk);
      } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + ooff + coff(ch);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float a0 = __uint_as_float(rr[16 * q + 2 * i]), a1 = __uint_as_float(rr[16 * q + 2 * i + 1]);
            const float off0 = EPI == EPI_MASK_LEAKY_BF16 ? kLeakySlope * a0 : 0.f, off1 = EPI == EPI_MASK_LEAKY_BF16 ? kLeakySlope * a1 : 0.f;
            const float lo = (mbits[ch] >> relu_bit_pos(16 * q + 2 * i)) & 1u ? a0 : off0;
            const float hi2 = (mbits[ch] >> relu_bit_pos(16 * q + 2 * i + 1)) & 1u ? a1 : off1;
            pk[i] = pack_bf16x2(lo, hi2);
            rr[16 * q + 2 * i] = __float_as_uint(lo);
            rr[16 * q + 2 * i + 1] = __float_as_uint(hi2);
          }
          stg_v8(o + 16 * q, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
        }
      }
    }
    if (epi_is_mask(EPI) && ok) {                     // fused bias gradient of the layer below
#pragma unroll
      for (int i = 0; i < 32; ++i) cs[(ch % BG) * 32 + i] += __uint_as_float(rr[i]);
    }
  }
}

// One CTA per SM: NR staged samples in flight, NACC accumulators in TMEM, NEPI epilogue warpgroups that take
// tiles round-robin.  The staging buffers are packed at the TMA box size: the 128-row MMA window of a tap
// may read past its sample into the next buffer (or the `slack_bytes` tail) — those rows belong to grid
// positions that are discarded, and a GEMM row only ever influences its own output row.
template <int N>
constexpr int tma_conv_smem(int KB, int raw_buf_bytes, int NR, int slack) {
  return KB * N * 128 + NR * raw_buf_bytes + slack + 1280 + 1024;
}

// KB / KBX / SHIFT_RW / PLANES (the tap geometry) are compile-time so the whole tap loop of a sample pair unrolls
// under ONE election and every descriptor is `base + immediate`.
template <int N, int EPI, int NR, int NACC, int NEPI, int MROWS, int KB, int KBX, int SHIFT_RW, int PLANES>
__global__ void __launch_bounds__(32 * (4 * NEPI + 2), 1) tma_conv_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                           const TmaConvArgs args) {
  using namespace umma;
  constexpr int B_BYTES = N * 128;
  constexpr int TMEM_COLS = NACC * N <= 64 ? 64 : NACC * N <= 128 ? 128 : NACC * N <= 256 ? 256 : 512;
  static_assert(NACC * N <= 512, "accumulators exceed tensor memory");
  constexpr int W_MMA = 4 * NEPI, W_TMA = 4 * NEPI + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                         // KB weight tiles, resident
  uint8_t* sRaw = smem + KB * B_BYTES;                        // NR staged samples
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + NR * args.raw_buf_bytes + args.slack_bytes);
  uint64_t* raw_full = bars;                    // [NR]
  uint64_t* raw_empty = bars + NR;              // [NR]
  uint64_t* tfull = bars + 2 * NR;              // [NACC]
  uint64_t* tempty = bars + 2 * NR + NACC;      // [NACC]
  uint64_t* b_full = bars + 2 * NR + 2 * NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NR + 2 * NACC + 1);
  static_assert((2 * NR + 2 * NACC + 2) * 8 <= 256, "barrier block overflows");
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (!epi_is_mask(EPI) && threadIdx.x < N) sBias[threadIdx.x] = args.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int a = 0; a < NR; ++a) { mbar_init(raw_full + a, 1); mbar_init(raw_empty + a, 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 4); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
  if (warp == W_TMA && lane == 0) tma_prefetch_desc(&tmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0) {
      // weights: resident for the whole kernel
      mbar_arrive_expect_tx(b_full, (uint32_t)(KB * B_BYTES));
      for (int kb = 0; kb < KB; ++kb)
        tma_bulk_g2s(sB + kb * B_BYTES, reinterpret_cast<const uint8_t*>(args.bpack) + (int64_t)kb * B_BYTES, B_BYTES, b_full);
      uint32_t j = 0;
      for (int b = blockIdx.x; b < args.nb; b += gridDim.x, ++j) {
        const uint32_t buf = j % NR;
        mbar_wait(raw_empty + buf, ((j / NR) & 1) ^ 1);
        mbar_arrive_expect_tx(raw_full + buf, (uint32_t)(args.box_bytes * PLANES));
        for (int pl = 0; pl < PLANES; ++pl)
          tma_load_4d(sRaw + buf * args.raw_buf_bytes + pl * args.plane_bytes, &tmap, 0, -args.pad, -args.pad + pl, b, raw_full + buf);
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    {
      constexpr uint32_t idesc = make_idesc_f16(1u, MROWS, N, 0, 0);
      constexpr uint32_t hi0 = desc_hi(1024, LAYOUT_SW128);
      const uint32_t b_lo0 = desc_lo(smem_u32(sB), 0);
      mbar_wait(b_full, 0);
      // Two samples are issued interleaved (k-block by k-block) into two accumulators: back-to-back MMAs into
      // ONE accumulator do not overlap their operand fetch with the previous MMA, two independent chains do.
      const int n_mine = (args.nb - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
      for (int j = 0; j < n_mine; j += 2) {
        const int np = min(2, n_mine - j);
        uint32_t raw_addr[2], tmem_d[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t jj = (uint32_t)(j + (t < np ? t : 0)), buf = jj % NR, acc = jj % NACC;
          mbar_wait(raw_full + buf, (jj / NR) & 1);
          mbar_wait(tempty + acc, ((jj / NACC) & 1) ^ 1);
          raw_addr[t] = smem_u32(sRaw + buf * args.raw_buf_bytes);
          tmem_d[t] = tmem_base + acc * N;
        }
        tc_fence_after();
        const uint32_t ab0 = desc_lo(raw_addr[0], 0), ab1 = desc_lo(raw_addr[1], 0), pl16 = (uint32_t)args.plane_bytes >> 4;
        auto issue = [&](auto two) {
          constexpr bool TWO = decltype(two)::value;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const int ty = kb / KBX, tx = kb % KBX;
            const uint32_t shift16 = PLANES == 2 ? (uint32_t)(((ty >> 1) * SHIFT_RW + tx) * 8) + ((ty & 1) ? pl16 : 0u)
                                                 : (uint32_t)((ty * SHIFT_RW + tx) * 8);
            const uint32_t b_lo = b_lo0 + kb * (B_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kb == 0 && k == 0) mma_f16_lohi<false>(tmem_d[0], ab0 + shift16, hi0, b_lo, hi0, idesc);
              else mma_f16_lohi<true>(tmem_d[0], ab0 + shift16 + 2 * k, hi0, b_lo + 2 * k, hi0, idesc);
            }
            if (TWO) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb == 0 && k == 0) mma_f16_lohi<false>(tmem_d[1], ab1 + shift16, hi0, b_lo, hi0, idesc);
                else mma_f16_lohi<true>(tmem_d[1], ab1 + shift16 + 2 * k, hi0, b_lo + 2 * k, hi0, idesc);
              }
            }
          }
          mma_commit(tfull + (uint32_t)j % NACC);
          mma_commit(raw_empty + (uint32_t)j % NR);       // staged sample free once every MMA reading it has retired
          if (TWO) {
            mma_commit(tfull + (uint32_t)(j + 1) % NACC);
            mma_commit(raw_empty + (uint32_t)(j + 1) % NR);
          }
        };
        if (elect_one()) {
          if (np == 2) issue(std::true_type{});
          else issue(std::false_type{});
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // fused bias gradient: every thread keeps running column sums of ITS rows (32 adds per chunk and tile);
    // the cross-lane reduction happens once, after the last tile.  Column chunk ch feeds bias group ch % BG.
    constexpr int BG = N == 64 ? 2 : 1;               // bias_mod / 32 (checked by the launcher)
    float cs[epi_is_mask(EPI) ? 32 * BG : 1];
#pragma unroll
    for (int i = 0; i < (epi_is_mask(EPI) ? 32 * BG : 1); ++i) cs[i] = 0.f;
    // MROWS = 64: the accumulator rows live in lanes 0-15 of each 32-lane quarter (row = 16*quarter + lane)
    const int wg = warp >> 2;
    const int trow = MROWS == 128 ? (int)(threadIdx.x & 127) : (lane < 16 ? 16 * (warp & 3) + lane : 1 << 20);
    uint32_t it = 0;
    for (int b = blockIdx.x; b < args.nb; b += gridDim.x) {
      for (int mt = 0; mt < args.MT; ++mt, ++it) {
        if (NEPI > 1 && (int)(it % NEPI) != wg) continue;
        const uint32_t acc = it % NACC;
        tma_conv_epilogue_tile<N, EPI, BG>(args, b, mt * 128 + trow, true, tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * N,
                                           sBias, cs, tfull + acc, (it / NACC) & 1, [&]() { mbar_arrive_warp(tempty + acc); });
      }
    }
    if (epi_is_mask(EPI) && args.bias_grad) {
#pragma unroll
      for (int g = 0; g < BG; ++g) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cs[g * 32 + i];
        atomicAdd(args.bias_grad + (g * 32 + lane) % args.bias_mod, warp_colsum32(v, lane));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace drl

namespace drl {

// ---------------------------------------------------------------------------------------------------------
// Weight gradient of a conv layer with TMA-staged operands and descriptor-shifted windows:
//   dW^T[kp = (ky,kx,ci)][co] += sum over positions r of  P[r][kp] * dY[r][co]
// Per sample the input activation (P source) and dY land in shared memory by tensor-map TMA loads; dY is
// laid out on the same width-RW row grid as the forward GEMM rows, its out-of-range columns/rows are zero
// (TMA out-of-bounds fill), so the discarded grid positions contribute nothing.  For tap t the MN-major A
// operand is the staged input shifted by the tap offset; two taps form one 128-row accumulator block
// (leading-dimension byte offset = distance between their windows).  All accumulator blocks stay in
// TMEM for the whole kernel; partial sums are added with red.global.add.f32 in OIHW order.
// ---------------------------------------------------------------------------------------------------------
struct WgradConvArgs {
  int nb;
  int ksteps;               // reduction extent in 16-row MMA steps
  int n_blocks;             // accumulator blocks (2 taps/atoms each)
  uint32_t a_off[8];        // byte offset of the first atom of each block inside the staged input
  uint32_t a_lbo[8];        // byte distance to the block's second atom
  int p_planes, p_pad, p_box_bytes, p_plane_bytes, p_buf_bytes;
  int dy_box_bytes, dy_buf_bytes;
  float* gw; int Cin, KH, KW, kp_valid;
  float* gt;                // when set: partial sums go to the [kp][NT] scratch with 16-byte reductions (see wgrad_untranspose_kernel)
  int pair_c;               // > 0: SAMPLE PAIRS packed along the channels (RND predictor): input channel = s * pair_c + ci, output
                            //    column = s' * 32 + co (NT = 64).  Only the diagonal blocks s' == s are gradients (the others are
                            //    products across the two samples): they are folded into the scratch row (tap, ci) of 32 columns.
};

template <int NT, int MB, int NR, int KSTEPS>
__global__ void __launch_bounds__(192, 1) wgrad_conv_tma_kernel(const __grid_constant__ CUtensorMap tmap_p,
                                                                const __grid_constant__ CUtensorMap tmap_dy,
                                                                const WgradConvArgs args) {
  using namespace umma;
  constexpr int ACC_COLS = MB * NT;
  constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stage_bytes = args.p_buf_bytes + args.dy_buf_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NR * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + NR;
  uint64_t* done = bars + 2 * NR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NR + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // rows the TMA never writes must read as zero (they meet zero dY rows, but 0 * NaN would poison the sum)
  for (int i = threadIdx.x * 16; i < NR * stage_bytes; i += blockDim.x * 16) st_shared_v4(smem_u32(smem + i), make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    for (int s = 0; s < NR; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool has_work = (int)blockIdx.x < args.nb;

  if (has_work) {
    if (warp == 5) {
      if (lane == 0) {
        uint32_t j = 0;
        for (int b = blockIdx.x; b < args.nb; b += gridDim.x, ++j) {
          const uint32_t buf = j % NR;
          mbar_wait(empty + buf, ((j / NR) & 1) ^ 1);
          mbar_arrive_expect_tx(full + buf, (uint32_t)(args.p_box_bytes * args.p_planes + args.dy_box_bytes));
          uint8_t* sp = smem + buf * stage_bytes;
          for (int pl = 0; pl < args.p_planes; ++pl)
            tma_load_4d(sp + pl * args.p_plane_bytes, &tmap_p, 0, -args.p_pad, -args.p_pad + pl, b, full + buf);
          tma_load_4d(sp + args.p_buf_bytes, &tmap_dy, 0, 0, 0, b, full + buf);
        }
      }
      __syncwarp();
    } else if (warp == 4) {
      {
        constexpr uint32_t idesc = make_idesc_f16(1u, 128, NT, 1, 1);       // bf16, MN-major A and B
        constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
        // descriptor low words of stage 0, per accumulator block: everything the loop adds is warp-uniform
        uint32_t a_base[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) a_base[mb] = desc_lo(smem_u32(smem) + args.a_off[mb], args.a_lbo[mb]);
        const uint32_t b_base = desc_lo(smem_u32(smem) + (uint32_t)args.p_buf_bytes, 0);
        const uint32_t stage16 = (uint32_t)stage_bytes >> 4;
        auto issue = [&](uint32_t so, auto first) {
          constexpr bool FIRST = decltype(first)::value;
#pragma unroll
          for (int mb = 0; mb < MB; ++mb)
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              const uint32_t al = a_base[mb] + so + k * ((16 * 128) >> 4), bl = b_base + so + k * ((16 * 128) >> 4);
              if (FIRST && k == 0) mma_f16_lohi<false>(tmem_base + mb * NT, al, hi, bl, hi, idesc);
              else mma_f16_lohi<true>(tmem_base + mb * NT, al, hi, bl, hi, idesc);
            }
        };
        uint32_t j = 0;
        for (int b = blockIdx.x; b < args.nb; b += gridDim.x, ++j) {
          const uint32_t buf = j % NR;
          mbar_wait(full + buf, (j / NR) & 1);
          tc_fence_after();
          if (elect_one()) {
            if (j == 0) issue(buf * stage16, std::true_type{});
            else issue(buf * stage16, std::false_type{});
            mma_commit(empty + buf);
          }
          __syncwarp();
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
      }
      __syncwarp();
    } else {
      mbar_wait(done, 0);
      tc_fence_after();
      const int khw = args.KH * args.KW;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int kp = mb * 128 + threadIdx.x;
        const bool ok = kp < args.kp_valid;
        const int kc = args.KW * args.Cin;
        const int ky = kp / kc, r2 = kp - ky * kc;
        const int kx = r2 / args.Cin, ci = r2 - kx * args.Cin;
        float* g0 = args.gw + ((int64_t)ci * args.KH + ky) * args.KW + kx;
#pragma unroll 1
        for (int ch = 0; ch < NT / 32; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * NT + ch * 32, r);
          tmem_ld_wait();
          if (ok && args.gt && args.pair_c) {
            const int cinp = 2 * args.pair_c, tap = kp / cinp, rem = kp - tap * cinp, s = rem / args.pair_c;
            if (ch == s) {
              float* t = args.gt + (int64_t)(tap * args.pair_c + rem - s * args.pair_c) * 32;
#pragma unroll
              for (int q = 0; q < 8; ++q)
                red_add_v4(t + 4 * q, __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                           __uint_as_float(r[4 * q + 3]));
            }
          } else if (ok && args.gt) {
            float* t = args.gt + (int64_t)kp * NT + ch * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              red_add_v4(t + 4 * q, __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                         __uint_as_float(r[4 * q + 3]));
          } else if (ok) {
#pragma unroll
            for (int jn = 0; jn < 32; ++jn) atomicAdd(g0 + (int64_t)(ch * 32 + jn) * args.Cin * khw, __uint_as_float(r[jn]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace drl
