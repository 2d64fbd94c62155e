// Fused policy/value heads + PPO losses + their backward (K5 + K6 + K8 + head part of K9).
//
// Replaces, for one minibatch: CategoricalPolicyHead logits (drl/agents/heads/policy_heads.py:107),
// SimpleValueHead (drl/agents/heads/value_heads.py:64), Categorical.log_prob/entropy
// (drl/algos/abstract.py:401-402), the three ppo_* loss functions and the composite
// (drl/algos/ppo.py:310-417, 195-200), and autograd's backward through all of them down to the
// pre-ReLU gradient of the 512 features (nature_cnn.py:35-36) plus the head weight/bias grads.
//
// One warp per sample: the 512 features are read once (coalesced), the A+K dot products are
// reduced with shuffles, every lane evaluates the scalar loss math, d(feat) is written once.
// Head weight gradients are accumulated in registers by the whole CTA over tiles of 8 samples
// (no shared-memory atomics) and flushed with one global atomicAdd per element per CTA.
#include "ppo_math.cuh"
#include "net.cuh"

namespace drl {

constexpr int kFeat = 512;
constexpr int kWarps = 8;


// a lane owns the column PAIRS (64q + 2*lane, +1), q = 0..7: features load as float2, d(feat) stores as one
// 4-byte bf16x2 (or float2) per pair
template <typename T> __device__ __forceinline__ void store_act2(T* p, float a, float b);
template <> __device__ __forceinline__ void store_act2<float>(float* p, float a, float b) {
  *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
template <> __device__ __forceinline__ void store_act2<__nv_bfloat16>(__nv_bfloat16* p, float a, float b) {
  *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
}

// MODE 0: forward only (logits/values/logp/entropy);  MODE 1: + losses;  MODE 2: + backward;
// MODE 3: backward of the heads from GIVEN d(logits) / d(values) (passed in logits_out / values_out): the
// vector-Jacobian product a differentiable Agent.forward needs (no loss is evaluated).
// OX = A + K when the launcher knows it at compile time (the Atari cases), else 0: the three O x 512 loops then
// run exactly O iterations; with OX = 0 their iterations o >= O are predicated off but still take issue slots.
// MINB = resident CTAs per SM the register allocation aims for: 2 (<= 128 registers; what the compiler chose on its own: 127)
// for the general instantiations, 3 (80 registers, 64 bytes of spill) for the exact-(A, K) training kernels with A <= 8: the
// kernel waits on its global loads (ncu: 16 resident warps, issue slots 49 % busy), so 24 resident warps pay: 0.094 -> 0.084 ms.
// (MINB = 1 lets the allocator take 168+ registers = one CTA per SM: 0.161 ms.)
template <int MAXA, int MODE, typename TD, int OX = 0, int MINB = 2>
__global__ void __launch_bounds__(kWarps * 32, MINB) heads_loss_kernel(
    const float* __restrict__ feat, HeadParams hp_ptrs, const int64_t* __restrict__ row_idx, int nb,
    drl_ppo_batch_t batch, drl_ppo_hyper_t hp, float inv_nb, float* __restrict__ losses_out,
    float* __restrict__ logits_out, float* __restrict__ values_out, float* __restrict__ logp_out,
    float* __restrict__ ent_out, TD* __restrict__ dpre) {
  constexpr int MAXO = MAXA + DRL_MAX_REWARD_STREAMS;
  constexpr int OLOOP = OX ? OX : MAXO;
  extern __shared__ float smem[];
  const int A = hp.num_actions, K = hp.num_streams, O = A + K;
  float* Ws = smem;                              // [O][512]
  float* bs = Ws + (size_t)O * kFeat;            // [O]
  float* sf = bs + MAXO;                         // [kWarps][512]  features of the tile
  float* sd = sf + kWarps * kFeat;               // [kWarps][MAXO] d(out) of the tile
  float* red = sd + kWarps * MAXO;               // [128]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int e = threadIdx.x; e < O * kFeat; e += blockDim.x) {
    const int o = e / kFeat, c = e % kFeat;
    Ws[e] = o < A ? hp_ptrs.wp[o * kFeat + c] : hp_ptrs.wv[o - A][c];
  }
  if (threadIdx.x < O) bs[threadIdx.x] = threadIdx.x < A ? hp_ptrs.bp[threadIdx.x] : hp_ptrs.bv[threadIdx.x - A][0];
  __syncthreads();

  float gw[MAXO][2];
  float gb = 0.f;
  float gfc[16];                                  // running column sums of this lane's d(feat) columns
#pragma unroll
  for (int q = 0; q < 16; ++q) gfc[q] = 0.f;
  if (MODE >= 2) {
#pragma unroll
    for (int o = 0; o < MAXO; ++o) { gw[o][0] = 0.f; gw[o][1] = 0.f; }
  }
  PpoSums acc{0.f, 0.f, 0.f, 0.f};
  const int ntiles = (nb + kWarps - 1) / kWarps;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int i = tile * kWarps + warp;
    const bool live = i < nb;
    float f[16];
    if (live) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float2 v = *reinterpret_cast<const float2*>(feat + (size_t)i * kFeat + 64 * q + 2 * lane);
        f[2 * q] = v.x; f[2 * q + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 16; ++q) f[q] = 0.f;
    }
    float out[MAXO];
#pragma unroll
    for (int o = OLOOP; o < MAXO; ++o) out[o] = 0.f;
#pragma unroll
    for (int o = 0; o < OLOOP; ++o) {
      if (OX || o < O) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float2 w = *reinterpret_cast<const float2*>(Ws + o * kFeat + 64 * q + 2 * lane);
          s = fmaf(f[2 * q], w.x, fmaf(f[2 * q + 1], w.y, s));
        }
        out[o] = warp_sum(s) + bs[o];
      } else {
        out[o] = 0.f;
      }
    }
    float logit[MAXA], dlogit[MAXA], v_new[DRL_MAX_REWARD_STREAMS], dv[DRL_MAX_REWARD_STREAMS];
#pragma unroll
    for (int j = 0; j < MAXA; ++j) logit[j] = j < A ? out[j] : 0.f;
#pragma unroll
    for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) {
      // out[A + k] with a compile-time-unrollable select
      float v = 0.f;
#pragma unroll
      for (int o = 0; o < MAXO; ++o) if (o == A + k) v = out[o];
      v_new[k] = v;
      dv[k] = 0.f;
    }
    if (MODE != 3 && live) {
      if (logits_out && lane < A) {
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < MAXA; ++j) if (j == lane) v = logit[j];
        logits_out[(size_t)i * A + lane] = v;
      }
      if (values_out && lane < K) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) if (k == lane) v = v_new[k];
        values_out[(size_t)i * K + lane] = v;
      }
    }
    if (MODE == 0) {
      if (live && (logp_out || ent_out)) {
        PpoSampleIn in{};
        in.action = batch.actions ? (int)batch.actions[row_idx ? row_idx[i] : (int64_t)i] : 0;
        PpoSums s;
        float logp, ent;
        drl_ppo_hyper_t h2 = hp;
        h2.num_streams = 0;
        ppo_sample<MAXA>(logit, v_new, in, h2, 0.f, false, dlogit, dv, s, logp, ent);
        if (lane == 0) {
          if (logp_out) logp_out[i] = logp;
          if (ent_out) ent_out[i] = ent;
        }
      }
      continue;
    }
    if (MODE == 3) {
      if (live) {                                  // upstream gradients instead of a loss
#pragma unroll
        for (int j = 0; j < MAXA; ++j) dlogit[j] = j < A ? logits_out[(size_t)i * A + j] : 0.f;
#pragma unroll
        for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) dv[k] = k < K ? values_out[(size_t)i * K + k] : 0.f;
      }
    } else if (live) {
      PpoSampleIn in;
      load_sample(batch, K, row_idx ? row_idx[i] : (int64_t)i, in);
      PpoSums s;
      float logp, ent;
      ppo_sample<MAXA, OX ? OX - 1 : 0, OX ? 1 : 0>(logit, v_new, in, hp, inv_nb, MODE == 2, dlogit, dv, s, logp, ent);   // OX is only dispatched for K = 1
      if (lane == 0) {
        acc.surr += s.surr; acc.ent += s.ent; acc.clip += s.clip; acc.vf += s.vf;
        if (logp_out) logp_out[i] = logp;
        if (ent_out) ent_out[i] = ent;
      }
    }
    if (MODE >= 2) {
      // d(out) of this sample -> smem; features -> smem (for the CTA-wide head wgrad)
      float dout[MAXO];
#pragma unroll
      for (int o = 0; o < MAXO; ++o) {
        float v = 0.f;
        if (live) {
#pragma unroll
          for (int j = 0; j < MAXA; ++j) if (j == o && o < A) v = dlogit[j];
#pragma unroll
          for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) if (o == A + k && k < K) v = dv[k];
        }
        dout[o] = v;
      }
      if (lane == 0) {
#pragma unroll
        for (int o = 0; o < MAXO; ++o) sd[warp * MAXO + o] = dout[o];
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float2*>(sf + warp * kFeat + 64 * q + 2 * lane) = make_float2(f[2 * q], f[2 * q + 1]);
      if (live) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float d0 = 0.f, d1 = 0.f;
#pragma unroll
          for (int o = 0; o < OLOOP; ++o) {
            if (OX || o < O) {
              const float2 w = *reinterpret_cast<const float2*>(Ws + o * kFeat + 64 * q + 2 * lane);
              d0 = fmaf(dout[o], w.x, d0);
              d1 = fmaf(dout[o], w.y, d1);
            }
          }
          d0 = f[2 * q] > 0.f ? d0 : 0.f;                 // ReLU of the fc layer (nature_cnn.py:36)
          d1 = f[2 * q + 1] > 0.f ? d1 : 0.f;
          gfc[2 * q] += d0;
          gfc[2 * q + 1] += d1;
          store_act2<TD>(dpre + (size_t)i * kFeat + 64 * q + 2 * lane, d0, d1);
        }
      }
      __syncthreads();
      const int c0 = threadIdx.x, c1 = threadIdx.x + kWarps * 32;
#pragma unroll
      for (int s = 0; s < kWarps; ++s) {
        const float f0 = sf[s * kFeat + c0], f1 = sf[s * kFeat + c1];
#pragma unroll
        for (int o = 0; o < OLOOP; ++o) {
          if (OX || o < O) {
            const float d = sd[s * MAXO + o];
            gw[o][0] = fmaf(d, f0, gw[o][0]);
            gw[o][1] = fmaf(d, f1, gw[o][1]);
          }
        }
        if (threadIdx.x < O) gb += sd[s * MAXO + threadIdx.x];
      }
      __syncthreads();
    }
  }
  if (MODE == 1 || MODE == 2) {
    PpoSums tot = block_reduce_sums(acc, red);
    if (threadIdx.x == 0) ppo_accumulate_losses(tot, hp, inv_nb, losses_out);
  }
  if (MODE >= 2) {
    const int c0 = threadIdx.x, c1 = threadIdx.x + kWarps * 32;
#pragma unroll
    for (int o = 0; o < MAXO; ++o) {
      if (o < O) {
        float* g = o < A ? hp_ptrs.gwp + o * kFeat : hp_ptrs.gwv[o - A];
        atomicAdd(g + c0, gw[o][0]);
        atomicAdd(g + c1, gw[o][1]);
      }
    }
    if (threadIdx.x < O) {
      float* g = threadIdx.x < A ? hp_ptrs.gbp + threadIdx.x : hp_ptrs.gbv[threadIdx.x - A];
      atomicAdd(g, gb);
    }
    if (hp_ptrs.gfc_bias) {                       // fc bias gradient: sum the 8 warps' column sums through smem
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float2*>(sf + warp * kFeat + 64 * q + 2 * lane) = make_float2(gfc[2 * q], gfc[2 * q + 1]);
      __syncthreads();
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) { t0 += sf[w * kFeat + c0]; t1 += sf[w * kFeat + c1]; }
      atomicAdd(hp_ptrs.gfc_bias + c0, t0);
      atomicAdd(hp_ptrs.gfc_bias + c1, t1);
    }
  }
}

extern int g_debug_flags[16];

template <int MAXA, int MODE, typename TD, int OX = 0>
static int launch_heads(const float* feat, const HeadParams& hpp, const int64_t* row_idx, int nb,
                        const drl_ppo_batch_t& batch, const drl_ppo_hyper_t& hp, float* losses_out,
                        float* logits_out, float* values_out, float* logp_out, float* ent_out,
                        TD* dpre, cudaStream_t st) {
  constexpr int MAXO = MAXA + DRL_MAX_REWARD_STREAMS;
  const int O = hp.num_actions + hp.num_streams;
  const size_t smem = sizeof(float) * ((size_t)O * kFeat + MAXO + kWarps * kFeat + kWarps * MAXO + 128);
  const int ntiles = (nb + kWarps - 1) / kWarps;
  if constexpr (OX != 0 && MAXA <= 8 && MODE == 2) {
    if (g_debug_flags[8]) {                              // 3 resident CTAs per SM (register-capped), one wave
      auto kern3 = heads_loss_kernel<MAXA, MODE, TD, OX, 3>;
      DRL_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int blocks3 = ntiles < sm_count() * 3 ? ntiles : sm_count() * 3;
      kern3<<<blocks3, kWarps * 32, smem, st>>>(feat, hpp, row_idx, nb, batch, hp, 1.0f / (float)nb, losses_out, logits_out,
                                                values_out, logp_out, ent_out, dpre);
      DRL_LAUNCH_CHECK();
      return DRL_OK;
    }
  }
  auto kern = heads_loss_kernel<MAXA, MODE, TD, OX>;
  DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int blocks = ntiles < sm_count() * 4 ? ntiles : sm_count() * 4;   // latency-bound: as many CTAs as the registers allow
  kern<<<blocks, kWarps * 32, smem, st>>>(feat, hpp, row_idx, nb, batch, hp, 1.0f / (float)nb,
                                          losses_out, logits_out, values_out, logp_out, ent_out, dpre);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// mode: 0 forward, 1 losses, 2 losses + backward, 3 backward from given d(logits) / d(values) (in logits_out /
// values_out).  dpre_is_bf16 selects the d(feat) dtype.
int heads_loss_dispatch(int mode, bool dpre_is_bf16, const float* feat, const HeadParams& hpp,
                        const int64_t* row_idx, int nb, const drl_ppo_batch_t& batch,
                        const drl_ppo_hyper_t& hp, float* losses_out, float* logits_out,
                        float* values_out, float* logp_out, float* ent_out, void* dpre,
                        cudaStream_t st) {
  const int A = hp.num_actions;
#define GO(MAXA, MODE, TD) \
  return launch_heads<MAXA, MODE, TD>(feat, hpp, row_idx, nb, batch, hp, losses_out, logits_out, \
                                      values_out, logp_out, ent_out, (TD*)dpre, st)
#define BY_MODE(MAXA)                                              \
  do {                                                             \
    if (mode == 0) GO(MAXA, 0, float);                             \
    if (mode == 1) GO(MAXA, 1, float);                             \
    if (mode == 3 && dpre_is_bf16) GO(MAXA, 3, __nv_bfloat16);     \
    if (mode == 3) GO(MAXA, 3, float);                             \
    if (dpre_is_bf16) GO(MAXA, 2, __nv_bfloat16);                  \
    GO(MAXA, 2, float);                                            \
  } while (0)
  // training step of the common Atari heads: exact output count known at compile time
#define GOX(MAXA, OXV)                                                                                              \
  do {                                                                                                              \
    if (dpre_is_bf16)                                                                                               \
      return launch_heads<MAXA, 2, __nv_bfloat16, OXV>(feat, hpp, row_idx, nb, batch, hp, losses_out, logits_out,    \
                                                       values_out, logp_out, ent_out, (__nv_bfloat16*)dpre, st);    \
    return launch_heads<MAXA, 2, float, OXV>(feat, hpp, row_idx, nb, batch, hp, losses_out, logits_out, values_out, \
                                             logp_out, ent_out, (float*)dpre, st);                                  \
  } while (0)
  if (mode == 2 && hp.num_streams == 1) {
    if (A == 4) GOX(8, 5);
    if (A == 6) GOX(8, 7);
    if (A == 9) GOX(18, 10);
    if (A == 18) GOX(18, 19);
  }
#undef GOX
  if (A <= 8) BY_MODE(8);
  if (A <= 18) BY_MODE(18);
  BY_MODE(32);
#undef BY_MODE
#undef GO
}

}  // namespace drl
