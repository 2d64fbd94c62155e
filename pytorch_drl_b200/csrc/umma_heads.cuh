// Policy / value heads + PPO losses + their backward on TENSOR CORES (the CUDA-core kernel of heads_loss.cu spent 0.087 ms per
// 32 768 samples on 3 x 7 x 512 dot products per sample: instruction-issue bound at 6x its HBM floor).
//
// Per tile of 128 samples (bf16 features [128 x 512] staged by TMA as 8 k-block tiles, 128B swizzle):
//   MMA1  logits / values   D1[128 x 32]  = feat x Wh^T          K-major A (features), K-major B (head weight tile images)
//   rows  one thread per sample: bias, Categorical log-prob / entropy, ppo_* losses and d(logits), d(values) (ppo_math.cuh)
//         -> G[128 x 32] bf16 into a swizzled shared-memory tile
//   MMA2  d(features)       D2[128 x 512] = G x Wh                K-major A (G), MN-major B — the SAME head-weight images
//         epilogue: ReLU mask of the fc layer from the staged feature tile, bf16 d(pre-activation) rows, fc bias gradient
//   MMA3  head weight grads D3[512 x 32] += feat^T x G            MN-major A (the staged feature tiles), MN-major B (G);
//         accumulated in tensor memory over the CTA's tiles, flushed once.
// One packed operand serves MMA1 and MMA2: a [head rows][64 features] swizzled tile is a K-major B operand with N = heads and,
// read along the other axis, an MN-major B operand with K = heads.
// Replaces CategoricalPolicyHead / SimpleValueHead (policy_heads.py:107, value_heads.py:64), Categorical.log_prob / entropy
// (abstract.py:401-402), ppo.py:310-417, 195-200 and their autograd backward, for A + K <= 32 on the bf16 engine.
#pragma once
#include <cuda.h>
#include "ppo_math.cuh"
#include "net.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"
#include "umma_conv1.cuh"
#include "umma_tma_gemm.cuh"

namespace drl {

constexpr int kHmNH = 32;                         // head columns of the MMAs (A + K padded)
constexpr int kHmWBytes = 8 * kHmNH * 128;        // head weight images: 8 k-blocks x [32 x 64] bf16
constexpr int kHmFBytes = 8 * 128 * 128;          // feature tile: 8 k-blocks x [128 x 64] bf16
constexpr int kHmGBytes = 128 * 128;
constexpr int kHmSmem = kHmWBytes + kHmFBytes + kHmGBytes + 512 * 4 + 256 + 256 + 1024;
constexpr int kHmThreads = 32 * (8 + 2);          // 2 epilogue warpgroups (warpgroup 0 also does the row math), MMA, TMA

struct HeadsMmaArgs {
  HeadParams hp_ptrs;
  const void* wpack;               // bf16 [8][32 x 64] swizzled head weight images (pack_all_kernel, PK_HEADS)
  const int64_t* row_idx;
  int nb;
  drl_ppo_batch_t batch;
  drl_ppo_hyper_t hp;
  float inv_nb;
  float* losses_out;
  __nv_bfloat16* dpre;             // [nb, 512]
};

// MODE 1: losses only (metrics pass); MODE 2: losses + backward.  AX / KX: exact action / stream counts or 0.
template <int MAXA, int MODE, int AX, int KX>
__global__ void __launch_bounds__(kHmThreads, 1) heads_mma_kernel(const __grid_constant__ CUtensorMap tmap_f, const HeadsMmaArgs args) {
  using namespace umma;
  constexpr int NH = kHmNH;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sW = smem;
  uint8_t* sF = sW + kHmWBytes;
  uint8_t* sG = sF + kHmFBytes;
  float* sFcB = reinterpret_cast<float*>(sG + kHmGBytes);            // [512] fc bias gradient of this CTA
  float* sBias = sFcB + 512;                                           // [32] head biases, then [32] head bias gradients
  float* sGb = sBias + 32;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sGb + 32);
  uint64_t* w_full = bars;          // head weights landed
  uint64_t* f_full = bars + 1;      // feature tile landed
  uint64_t* f_free = bars + 2;      // MMAs that read the feature tile / G retired (commit)
  uint64_t* e_done = bars + 3;      // epilogue threads done with the feature tile (ReLU masks): 256 arrivals
  uint64_t* d1_full = bars + 4;
  uint64_t* d1_free = bars + 5;     // 128 arrivals
  uint64_t* g_full = bars + 6;      // 128 arrivals
  uint64_t* d2_full = bars + 7;     // [2]
  uint64_t* d2_free = bars + 9;     // [2] 128 arrivals each
  uint64_t* d3_full = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* red = reinterpret_cast<float*>(bars + 14);                    // [128] loss reduction scratch (inside the 256 + 256 block)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int A = AX ? AX : args.hp.num_actions, K = KX ? KX : args.hp.num_streams, O = A + K;
  const int ntiles = (args.nb + 127) / 128;
  const int n_mine = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  for (int i = threadIdx.x; i < 512 + 64; i += blockDim.x) sFcB[i] = 0.f;       // fc bias gradient, biases, head bias gradients
  if (threadIdx.x == 0) {
    mbar_init(w_full, 1); mbar_init(f_full, 1); mbar_init(f_free, 1); mbar_init(e_done, 256);
    mbar_init(d1_full, 1); mbar_init(d1_free, 128); mbar_init(g_full, 128);
    for (int a = 0; a < 2; ++a) { mbar_init(d2_full + a, 1); mbar_init(d2_free + a, 128); }
    mbar_init(d3_full, 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  if (warp == 9 && lane == 0) tma_prefetch_desc(&tmap_f);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x < O) sBias[threadIdx.x] = threadIdx.x < A ? args.hp_ptrs.bp[threadIdx.x] : args.hp_ptrs.bv[threadIdx.x - A][0];
  __syncthreads();                                   // biases visible to the row threads
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t D1 = 0, D3 = 64, D2 = 256;     // tensor-memory columns: logits | 4 x 32 head weight gradient blocks | 2 x 128 d(feat)

  if (warp == 9) {
    if (lane == 0 && n_mine > 0) {
      mbar_arrive_expect_tx(w_full, kHmWBytes);
      tma_bulk_g2s(sW, args.wpack, kHmWBytes, w_full);
      for (int j = 0; j < n_mine; ++j) {
        const int tile = blockIdx.x + j * gridDim.x;
        if (j > 0) {                                 // the previous tile's readers (MMA2 / MMA3 and the mask reads) are done
          mbar_wait(f_free, (j - 1) & 1);
          if (MODE == 2) mbar_wait(e_done, (j - 1) & 1);
        }
        mbar_arrive_expect_tx(f_full, kHmFBytes);
        for (int kb = 0; kb < 8; ++kb) tma_load_2d(sF + kb * 16384, &tmap_f, kb * 64, tile * 128, f_full);
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
    constexpr uint32_t id1 = make_idesc_f16(1u, 128, NH, 0, 0);       // logits: both K-major
    constexpr uint32_t id2 = make_idesc_f16(1u, 128, 128, 0, 1);      // d(feat): A = G K-major, B = head images MN-major
    constexpr uint32_t id3 = make_idesc_f16(1u, 128, NH, 1, 1);       // head wgrad: both MN-major
    const uint32_t f_lo = desc_lo(smem_u32(sF), 0), w_lo = desc_lo(smem_u32(sW), 0), g_lo = desc_lo(smem_u32(sG), 0);
    const uint32_t w_mn = desc_lo(smem_u32(sW), NH * 128);            // MN-major: the next 64-feature atom is the next image
    const uint32_t f_mn = desc_lo(smem_u32(sF), 16384);               // MN-major M = 128 features = two k-block tiles
    if (n_mine > 0) mbar_wait(w_full, 0);
    for (int j = 0; j < n_mine; ++j) {
      mbar_wait(f_full, j & 1);
      if (j > 0) mbar_wait(d1_free, (j - 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 8; ++kb)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (kb == 0 && k == 0) mma_f16_lohi<false>(tmem_base + D1, f_lo, hi, w_lo, hi, id1);
            else mma_f16_lohi<true>(tmem_base + D1, f_lo + kb * (16384 >> 4) + 2 * k, hi, w_lo + kb * ((NH * 128) >> 4) + 2 * k, hi, id1);
          }
        mma_commit(d1_full);
        if (MODE == 1) mma_commit(f_free);
      }
      __syncwarp();
      if (MODE == 2) {
        mbar_wait(g_full, j & 1);
        tc_fence_after();
        // d(feat) in four 128-column chunks, two accumulator buffers (one per epilogue warpgroup)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const uint32_t use = (uint32_t)(2 * j + (c >> 1));           // how often buffer c & 1 has been used before
          mbar_wait(d2_free + (c & 1), (use & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < NH / 16; ++ks)
              if (ks == 0) mma_f16_lohi<false>(tmem_base + D2 + (c & 1) * 128, g_lo, hi, w_mn + 2 * c * ((NH * 128) >> 4), hi, id2);
              else mma_f16_lohi<true>(tmem_base + D2 + (c & 1) * 128, g_lo + 2 * ks, hi, w_mn + 2 * c * ((NH * 128) >> 4) + ks * 128, hi, id2);
            mma_commit(d2_full + (c & 1));
          }
          __syncwarp();
        }
        if (elect_one()) {
#pragma unroll
          for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (j == 0 && ks == 0) mma_f16_lohi<false>(tmem_base + D3 + b * NH, f_mn + 2 * b * (16384 >> 4), hi, g_lo, hi, id3);
              else mma_f16_lohi<true>(tmem_base + D3 + b * NH, f_mn + 2 * b * (16384 >> 4) + ks * 128, hi, g_lo + ks * 128, hi, id3);
            }
          mma_commit(f_free);                       // feature tile and G consumed by every MMA of this tile
          if (j == n_mine - 1) mma_commit(d3_full);
        }
        __syncwarp();
      }
    }
  } else {
    const int wg = warp >> 2, row = threadIdx.x & 127;
    const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
    PpoSums acc{0.f, 0.f, 0.f, 0.f};
    float gb[MAXA + DRL_MAX_REWARD_STREAMS];          // head bias gradients of this thread's samples
#pragma unroll
    for (int o = 0; o < MAXA + DRL_MAX_REWARD_STREAMS; ++o) gb[o] = 0.f;
    for (int j = 0; j < n_mine; ++j) {
      const int tile = blockIdx.x + j * gridDim.x;
      const int i = tile * 128 + row;
      const bool live = i < args.nb;
      if (wg == 0) {
        // ---- row math: one thread per sample ----
        PpoSampleIn in{};
        if (live) load_sample(args.batch, K, args.row_idx ? args.row_idx[i] : (int64_t)i, in);      // global loads in flight during MMA1
        mbar_wait(d1_full, j & 1);
        tc_fence_after();
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + t_lane + D1, r);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(d1_free);
        float logit[MAXA], dlogit[MAXA], v_new[DRL_MAX_REWARD_STREAMS], dv[DRL_MAX_REWARD_STREAMS];
#pragma unroll
        for (int a = 0; a < MAXA; ++a) { logit[a] = a < A ? __uint_as_float(r[a < 32 ? a : 0]) + sBias[a < 32 ? a : 0] : 0.f; dlogit[a] = 0.f; }
#pragma unroll
        for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) {
          float v = 0.f;
#pragma unroll
          for (int o = 0; o < 32; ++o) if (o == A + k && k < K) v = __uint_as_float(r[o]) + sBias[o];
          v_new[k] = v; dv[k] = 0.f;
        }
        if (live) {
          PpoSums s;
          float logp, ent;
          ppo_sample<MAXA, AX, KX>(logit, v_new, in, args.hp, args.inv_nb, MODE == 2, dlogit, dv, s, logp, ent);
          acc.surr += s.surr; acc.ent += s.ent; acc.clip += s.clip; acc.vf += s.vf;
        }
        if (MODE == 2) {
          // G row: d(logits) | d(values) | zeros as 32 bf16 = chunks 0-3 of the swizzled 128-byte row
          float dout[32];
#pragma unroll
          for (int o = 0; o < 32; ++o) {
            float v = 0.f;
            if (live) {
#pragma unroll
              for (int a = 0; a < MAXA; ++a) if (a == o && a < A) v = dlogit[a];
#pragma unroll
              for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) if (o == A + k && k < K) v = dv[k];
            }
            dout[o] = v;
            if (o < MAXA + DRL_MAX_REWARD_STREAMS) gb[o] += v;
          }
          if (j > 0) mbar_wait(f_free, (j - 1) & 1);                   // the previous tile's MMAs are done reading G
          const uint32_t grow = smem_u32(sG) + (uint32_t)row * 128u;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            st_shared_v4(grow + (((uint32_t)c ^ ((uint32_t)row & 7u)) << 4),
                         make_uint4(pack_bf16x2(dout[8 * c], dout[8 * c + 1]), pack_bf16x2(dout[8 * c + 2], dout[8 * c + 3]),
                                    pack_bf16x2(dout[8 * c + 4], dout[8 * c + 5]), pack_bf16x2(dout[8 * c + 6], dout[8 * c + 7])));
          fence_proxy_async_smem();
          mbar_arrive(g_full);
        }
      }
      if (MODE == 2) {
        // ---- d(feat) epilogue: warpgroup wg drains accumulator buffer wg: column chunks c = wg and wg + 2 ----
        mbar_wait(f_full, j & 1);                    // (the mask reads below use the staged feature tile)
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = wg + 2 * cc;
          const uint32_t use = (uint32_t)(2 * j + cc);
          mbar_wait(d2_full + wg, use & 1);
          tc_fence_after();
#pragma unroll 1
          for (int sub = 0; sub < 4; ++sub) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tmem_base + t_lane + D2 + wg * 128 + sub * 32, r);
            tmem_ld_wait();
            if (sub == 3) {
              tc_fence_before();
              mbar_arrive(d2_free + wg);
            }
            // ReLU mask of the fc layer: features [128 c + 32 sub, + 32) of this row from the staged tile (k-block 2c + sub/2)
            const int col0 = 128 * c + 32 * sub;
            const uint32_t frow = smem_u32(sF) + (uint32_t)((col0 >> 6) * 16384 + row * 128);
            uint32_t fw[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 v = lds_v4(frow + ((((uint32_t)((col0 & 63) >> 3) + q) ^ ((uint32_t)row & 7u)) << 4));
              fw[4 * q] = v.x; fw[4 * q + 1] = v.y; fw[4 * q + 2] = v.z; fw[4 * q + 3] = v.w;
            }
            const uint32_t mb = relu_bits_from_pairs(fw);
            float vals[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) vals[e] = ((mb >> relu_bit_pos(e)) & 1u) && live ? __uint_as_float(r[e]) : 0.f;
            if (live) {
              __nv_bfloat16* o = args.dpre + (int64_t)i * 512 + col0;
#pragma unroll
              for (int q = 0; q < 2; ++q)
                stg_v8(o + 16 * q, pack_bf16x2(vals[16 * q], vals[16 * q + 1]), pack_bf16x2(vals[16 * q + 2], vals[16 * q + 3]),
                       pack_bf16x2(vals[16 * q + 4], vals[16 * q + 5]), pack_bf16x2(vals[16 * q + 6], vals[16 * q + 7]),
                       pack_bf16x2(vals[16 * q + 8], vals[16 * q + 9]), pack_bf16x2(vals[16 * q + 10], vals[16 * q + 11]),
                       pack_bf16x2(vals[16 * q + 12], vals[16 * q + 13]), pack_bf16x2(vals[16 * q + 14], vals[16 * q + 15]));
            }
            if (args.hp_ptrs.gfc_bias) atomicAdd(sFcB + col0 + lane, warp_colsum32(vals, lane));   // fc bias gradient (shared-memory atomics)
          }
        }
        mbar_arrive(e_done);
      }
    }
    // ---- after the last tile: losses, bias gradients, head weight gradients ----
    if (wg == 0) {
      acc.surr = warp_sum(acc.surr); acc.ent = warp_sum(acc.ent); acc.clip = warp_sum(acc.clip); acc.vf = warp_sum(acc.vf);
      if (lane == 0) { red[warp] = acc.surr; red[4 + warp] = acc.ent; red[8 + warp] = acc.clip; red[12 + warp] = acc.vf; }
      if (MODE == 2) {
#pragma unroll
        for (int o = 0; o < MAXA + DRL_MAX_REWARD_STREAMS; ++o) {
          const float s = warp_sum(gb[o]);
          if (lane == 0 && o < O) atomicAdd(sGb + o, s);
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 0 && n_mine > 0) {
      PpoSums tot{red[0] + red[1] + red[2] + red[3], red[4] + red[5] + red[6] + red[7], red[8] + red[9] + red[10] + red[11],
                  red[12] + red[13] + red[14] + red[15]};
      ppo_accumulate_losses(tot, args.hp, args.inv_nb, args.losses_out);
    }
    if (MODE == 2 && n_mine > 0) {
      if (threadIdx.x < O) {
        float* g = threadIdx.x < A ? args.hp_ptrs.gbp + threadIdx.x : args.hp_ptrs.gbv[threadIdx.x - A];
        atomicAdd(g, sGb[threadIdx.x]);
      }
      if (args.hp_ptrs.gfc_bias)
        for (int c = threadIdx.x; c < 512; c += 256) atomicAdd(args.hp_ptrs.gfc_bias + c, sFcB[c]);
      if (wg == 0) {
        mbar_wait(d3_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + t_lane + D3 + b * NH, r);
          tmem_ld_wait();
          const int feat_col = 128 * b + row;
#pragma unroll
          for (int o = 0; o < 32; ++o) {
            if (o < O) {
              float* g = o < A ? args.hp_ptrs.gwp + (int64_t)o * 512 : args.hp_ptrs.gwv[o - A];
              atomicAdd(g + feat_col, __uint_as_float(r[o]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace drl

namespace drl {

// tensor map helper lives in umma_launch.cuh (make_tmap_bf16); declared here to keep this header self-contained for net.cu
int heads_mma_launch(int mode, const void* featb, const void* wpack, const HeadParams& hpp, const int64_t* row_idx, int nb,
                     const drl_ppo_batch_t& batch, const drl_ppo_hyper_t& hp, float* losses_out, void* dpre, cudaStream_t st);

}  // namespace drl
