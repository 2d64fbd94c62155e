// Family 3 — plain GEMMs of the fc layer with BOTH operands delivered by the TMA engine
// (cp.async.bulk.tensor.2d through CUtensorMap descriptors with 128B swizzle, SASS: UTMALDG):
// no producer warps, no generic-proxy writes, no proxy fences.
//
//   tma_gemm_kernel<BN, EPI>   D[128 x BN] = A[128 x K] * B[BN x K]^T   (both K-major)
//       fc forward  : A = act3 [nb, 3136] bf16, B = Wfc (NHWC-permuted) [512, 3136]  -> feat fp32 (+bias, ReLU)
//       fc data grad: A = dpre [nb, 512]  bf16, B = Wfc^T               [3136, 512]  -> dact3 bf16 (ReLU mask)
//   wgrad_tma_kernel<NT, MB>   dW^T[128*MB x NT] += P[m x .]^T * dY[m x NT]          (both MN-major)
//       fc weight gradient: P = act3 [nb, 3136], dY = dpre [nb, 512]
//
// Warp roles: 0-3 epilogue (TMEM lanes 32*warp), 4 MMA issuer, 5 TMA issuer.  Persistent CTAs.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace drl {

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tmap), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

struct TmaGemmArgs {
  int M, N_total, KB;          // rows, output columns, 64-element k-blocks
  void* out;                   // fp32 or bf16, row pitch out_ld elements
  int out_ld;
  const float* bias;           // EPI_BIAS_RELU_F32
  const __nv_bfloat16* mask_src;   // EPI_MASK_BF16 (same layout as out)
  const uint32_t* bits_in;     // EPI_MASK_BF16: ReLU mask words (umma.cuh relu_bit_pos) of act, element e = m*out_ld + n
  float* bias_grad;            // EPI_MASK_*: += column sums of the produced gradient at index column % bias_mod (64 or 32)
  int bias_mod;
  int pair_out;                // EPI_MASK_LEAKY_BF16 (RND fc1 data gradient): row m, 32-column chunk c of the [M, out_ld] result goes to
                               //   the SAMPLE-PAIR packed map  out[(m >> 1) * 2 * out_ld + c * 64 + (m & 1) * 32]  ([pairs][pixel c][2][32]);
                               //   the mask bits stay in the plain [M, out_ld] order; an odd M gets a zero phantom row M.
};

template <int BN>
struct TmaGemmCfg {
  static constexpr int kStages = 4;
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStage = kABytes + kBBytes;
  static constexpr int kSmem = kStages * kStage + 256 + 1024 + 1024;   // stages, barriers, bias, alignment slack
};

// Warps 0-7: two epilogue warpgroups, each drains HALF the columns of every tile (the epilogue of a 128 x 256 tile
// costs about as much as its 32 MMAs, so one warpgroup was the critical path); warp 8: MMA; warp 9: TMA.
// NEPI = 4 (data gradient: mask + bf16 pack + column sums per chunk): four warpgroups, chunk ch = wg + 4 * c2, so a
// thread only ever sees ONE of the two 32-column bias groups (32 running sums instead of 64: fits 112 registers).
constexpr int kTmaGemmThreads = 320;
template <int BN, int EPI, int NEPI = 2>
__global__ void __launch_bounds__(32 * (4 * NEPI + 2), 1) tma_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                          const __grid_constant__ CUtensorMap tmap_b,
                                                          const TmaGemmArgs args) {
  using namespace umma;
  using Cf = TmaGemmCfg<BN>;
  constexpr int S = Cf::kStages;
  constexpr int TMEM_COLS = 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cf::kStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (args.M + 127) / 128, n_tiles = (args.N_total + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 4 * NEPI); }
    fence_barrier_init();
  }
  constexpr int W_MMA = 4 * NEPI, W_TMA = 4 * NEPI + 1;
  static_assert(NEPI == 2 || NEPI == 4, "epilogue warpgroups");
  if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
  if (warp == W_TMA && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile % m_tiles, nt = tile / m_tiles;
        for (int kb = 0; kb < args.KB; ++kb, ++it) {
          const uint32_t stage = it % S, phase = (it / S) & 1;
          mbar_wait(empty + stage, phase ^ 1);
          mbar_arrive_expect_tx(full + stage, Cf::kStage);
          uint8_t* st = smem + stage * Cf::kStage;
          tma_load_2d(st, &tmap_a, kb * 64, mt * 128, full + stage);
          tma_load_2d(st + Cf::kABytes, &tmap_b, kb * 64, nt * BN, full + stage);
        }
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    {
      constexpr uint32_t idesc = make_idesc_f16(1u, 128, BN, 0, 0);
      constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
      const uint32_t a_lo0 = desc_lo(smem_u32(smem), 0), b_lo0 = desc_lo(smem_u32(smem) + Cf::kABytes, 0);
      uint32_t it = 0, j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
        const uint32_t acc = j & 1;
        mbar_wait(tempty + acc, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < args.KB; ++kb, ++it) {
          const uint32_t stage = it % S, phase = (it / S) & 1;
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (Cf::kStage >> 4), b_lo = b_lo0 + stage * (Cf::kStage >> 4);
          if (elect_one()) {
            if (kb == 0) mma_f16_lohi<false>(tmem_d, a_lo, hi, b_lo, hi, idesc);
            else mma_f16_lohi<true>(tmem_d, a_lo, hi, b_lo, hi, idesc);
            mma_f16_lohi<true>(tmem_d, a_lo + 2, hi, b_lo + 2, hi, idesc);
            mma_f16_lohi<true>(tmem_d, a_lo + 4, hi, b_lo + 4, hi, idesc);
            mma_f16_lohi<true>(tmem_d, a_lo + 6, hi, b_lo + 6, hi, idesc);
            mma_commit(empty + stage);
            if (kb == args.KB - 1) mma_commit(tfull + acc);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    uint32_t j = 0;
    int last_nt = -1;
    // fused bias gradient: per-thread running column sums (bias index = column % 64, two 32-column groups);
    // one cross-lane reduction after the last tile
    constexpr int BG = NEPI == 2 ? 2 : 1;                 // bias groups a thread meets
    constexpr int CPW = BN / 32 / NEPI;                   // 32-column chunks per warpgroup and tile
    float cs[epi_is_mask(EPI) ? 32 * BG : 1];
#pragma unroll
    for (int i = 0; i < (epi_is_mask(EPI) ? 32 * BG : 1); ++i) cs[i] = 0.f;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const int mt = tile % m_tiles, nt = tile / m_tiles;
      const uint32_t acc = j & 1;
      const int m = mt * 128 + (int)(threadIdx.x & 127);
      const int n0 = nt * BN;
      if (!epi_is_mask(EPI) && nt != last_nt) {            // (re)load the bias slice of this N tile
        asm volatile("bar.sync 1, %0;" ::"n"(128 * NEPI));
        for (int i = threadIdx.x; i < BN; i += 128 * NEPI) sBias[i] = n0 + i < args.N_total ? args.bias[n0 + i] : 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(128 * NEPI));
        last_nt = nt;
      }
      mbar_wait(tfull + acc, (j >> 1) & 1);
      tc_fence_after();
      const int wg = warp >> 2;                             // NEPI = 2: warpgroup 0 takes columns [0, BN/2), 1 the rest
#pragma unroll
      for (int c2 = 0; c2 < CPW; ++c2) {
        const int ch = NEPI == 2 ? wg * CPW + c2 : wg + NEPI * c2;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * BN + ch * 32, r);
        tmem_ld_wait();
        if (c2 == CPW - 1) {
          tc_fence_before();
          mbar_arrive_warp(tempty + acc);
        }
        const int n = n0 + ch * 32;
        if (m < args.M && n < args.N_total) {
          if (EPI == EPI_BIAS_RELU_F32 || EPI == EPI_BIAS_F32) {
            float* o = reinterpret_cast<float*>(args.out) + (int64_t)m * args.out_ld + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float x = __uint_as_float(r[8 * q + i]) + sBias[ch * 32 + 8 * q + i];
                v[i] = __float_as_uint(EPI == EPI_BIAS_RELU_F32 ? fmaxf(x, 0.f) : x);
              }
              stg_v8(o + 8 * q, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            }
          } else if (EPI == EPI_BIAS_RELU_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + (int64_t)m * args.out_ld + n;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                pk[i] = pack_bf16x2_relu(__uint_as_float(r[16 * q + 2 * i]) + sBias[ch * 32 + 16 * q + 2 * i],
                                         __uint_as_float(r[16 * q + 2 * i + 1]) + sBias[ch * 32 + 16 * q + 2 * i + 1]);
              stg_v8(o + 16 * q, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            }
          } else {
            const int64_t off = (int64_t)m * args.out_ld + n;
            const int64_t ooff = (EPI == EPI_MASK_LEAKY_BF16 && args.pair_out)
                                     ? (int64_t)(m >> 1) * 2 * args.out_ld + (int64_t)(n >> 5) * 64 + (m & 1) * 32 : off;
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + ooff;
            uint32_t mb;                                   // activation-derivative mask of these 32 columns, 1 bit per element
            if (args.bits_in) {
              mb = __ldg(args.bits_in + (off >> 5));
            } else {                                       // fallback: derive it from the bf16 activation itself
              uint32_t mw[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 mv = ldg_v4(args.mask_src + off + 8 * q);
                mw[4 * q] = mv.x; mw[4 * q + 1] = mv.y; mw[4 * q + 2] = mv.z; mw[4 * q + 3] = mv.w;
              }
              mb = EPI == EPI_MASK_LEAKY_BF16 ? pos_bits_from_pairs(mw) : relu_bits_from_pairs(mw);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float a0 = __uint_as_float(r[16 * q + 2 * i]), a1 = __uint_as_float(r[16 * q + 2 * i + 1]);
                const float off0 = EPI == EPI_MASK_LEAKY_BF16 ? kLeakySlope * a0 : 0.f, off1 = EPI == EPI_MASK_LEAKY_BF16 ? kLeakySlope * a1 : 0.f;
                const float lo = (mb >> relu_bit_pos(16 * q + 2 * i)) & 1u ? a0 : off0;
                const float hi2 = (mb >> relu_bit_pos(16 * q + 2 * i + 1)) & 1u ? a1 : off1;
                pk[i] = pack_bf16x2(lo, hi2);
                r[16 * q + 2 * i] = __float_as_uint(lo);
                r[16 * q + 2 * i + 1] = __float_as_uint(hi2);
              }
              stg_v8(o + 16 * q, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            }
          }
        } else if (EPI == EPI_MASK_LEAKY_BF16 && args.pair_out && m == args.M && (args.M & 1) && n < args.N_total) {
          // odd number of samples: the second half of the last pair is a phantom sample whose gradient must read as zero
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + (int64_t)(m >> 1) * 2 * args.out_ld + (int64_t)(n >> 5) * 64 + 32;
          stg_v8(o, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
          stg_v8(o + 16, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u);
        }
        if (epi_is_mask(EPI) && m < args.M && n < args.N_total) {       // fused bias gradient of the layer below
#pragma unroll
          for (int i = 0; i < 32; ++i) cs[(c2 % BG) * 32 + i] += __uint_as_float(r[i]);   // NEPI = 2: ch = wg * CPW + c2, CPW even; NEPI = 4: ch % 2 = wg % 2
        }
      }
    }
    if (epi_is_mask(EPI) && args.bias_grad) {
#pragma unroll
      for (int g = 0; g < BG; ++g) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cs[g * 32 + i];
        atomicAdd(args.bias_grad + ((NEPI == 2 ? g : ((warp >> 2) & 1)) * 32 + lane) % args.bias_mod, warp_colsum32(v, lane));
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

// ---- weight gradient with TMA-fed MN-major operands -------------------------------------------------------
struct WgradTmaArgs {
  int M;                 // reduction rows (samples)
  int n_tiles;           // N tiles of NT columns
  int kp_valid;          // valid accumulator rows (3136)
  float* gw;             // OIHW fp32 gradient, kp = (ky, kx, ci)
  int Cin, KH, KW;
  float* gt; int gt_ld;  // when set: partial sums go to the [kp][gt_ld] scratch with 16-byte reductions
};

template <int NT, int MB>
struct WgradTmaCfg {
  static constexpr int kRows = 32, kStages = 4;
  static constexpr int kAtom = kRows * 128;                  // 32 rows x 64 elements
  static constexpr int kPBytes = 2 * MB * kAtom;
  static constexpr int kDyBytes = (NT / 64) * kAtom;
  static constexpr int kStage = kPBytes + kDyBytes;
  static constexpr int kSmem = kStages * kStage + 256 + 1024;
};

template <int NT, int MB>
__global__ void __launch_bounds__(192, 2) wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmap_p,
                                                           const __grid_constant__ CUtensorMap tmap_dy,
                                                           const WgradTmaArgs args) {
  using namespace umma;
  using Cf = WgradTmaCfg<NT, MB>;
  constexpr int S = Cf::kStages, ROWS = Cf::kRows, NA = 2 * MB;
  constexpr int ACC_COLS = MB * NT;
  constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cf::kStage);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* done = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_stages = (args.M + ROWS - 1) / ROWS;
  const int per = (total_stages + gridDim.x - 1) / gridDim.x;
  const int st_begin = blockIdx.x * per, st_end = min(total_stages, st_begin + per);
  const int n_tile = blockIdx.y % args.n_tiles, mb_group = blockIdx.y / args.n_tiles;
  const int kp0 = mb_group * NA * 64, n0 = n_tile * NT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (st_begin < st_end) {
    if (warp == 5) {
      if (lane == 0) {
        uint32_t it = 0;
        for (int st = st_begin; st < st_end; ++st, ++it) {
          const uint32_t stage = it % S, phase = (it / S) & 1;
          mbar_wait(empty + stage, phase ^ 1);
          mbar_arrive_expect_tx(full + stage, Cf::kStage);
          uint8_t* sp = smem + stage * Cf::kStage;
#pragma unroll
          for (int a = 0; a < NA; ++a) tma_load_2d(sp + a * Cf::kAtom, &tmap_p, kp0 + a * 64, st * ROWS, full + stage);
#pragma unroll
          for (int a = 0; a < NT / 64; ++a)
            tma_load_2d(sp + Cf::kPBytes + a * Cf::kAtom, &tmap_dy, n0 + a * 64, st * ROWS, full + stage);
        }
      }
      __syncwarp();
    } else if (warp == 4) {
      {
        constexpr uint32_t idesc = make_idesc_f16(1u, 128, NT, 1, 1);
        constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
        const uint32_t a_lo0 = desc_lo(smem_u32(smem), Cf::kAtom);
        const uint32_t b_lo0 = desc_lo(smem_u32(smem) + Cf::kPBytes, Cf::kAtom);
        uint32_t it = 0;
        for (int st = st_begin; st < st_end; ++st, ++it) {
          const uint32_t stage = it % S, phase = (it / S) & 1;
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (Cf::kStage >> 4), b_lo = b_lo0 + stage * (Cf::kStage >> 4);
          if (elect_one()) {
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
              for (int k = 0; k < ROWS / 16; ++k) {
                const uint32_t al = a_lo + mb * ((2 * Cf::kAtom) >> 4) + k * ((16 * 128) >> 4);
                const uint32_t bl = b_lo + k * ((16 * 128) >> 4);
                if (k == 0 && it == 0) mma_f16_lohi<false>(tmem_base + mb * NT, al, hi, bl, hi, idesc);
                else mma_f16_lohi<true>(tmem_base + mb * NT, al, hi, bl, hi, idesc);
              }
            mma_commit(empty + stage);
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
        const int kp = kp0 + mb * 128 + threadIdx.x;
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
          if (ok && args.gt) {
            float* t = args.gt + (int64_t)kp * args.gt_ld + n0 + ch * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              red_add_v4(t + 4 * q, __uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                         __uint_as_float(r[4 * q + 3]));
          } else if (ok) {
#pragma unroll
            for (int jn = 0; jn < 32; ++jn)
              atomicAdd(g0 + (int64_t)(n0 + ch * 32 + jn) * args.Cin * khw, __uint_as_float(r[jn]));
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
