// Per-sample PPO loss math shared by the stand-alone fused loss kernel (ppo_loss.cu) and the
// fused heads+loss+backward kernel (heads_loss.cu).
//
// Restates Categorical.log_prob / entropy (torch; called at drl/algos/abstract.py:401-402),
// ppo_policy_entropy_bonus (drl/algos/ppo.py:310-328), ppo_policy_surrogate_objective
// (ppo.py:331-367), ppo_vf_loss (ppo.py:370-417) and their autograd derivatives, including
// torch's tie rules (min/max split the gradient evenly on ties; clamp passes it on the closed
// interval).
#pragma once
#include "common.cuh"

namespace drl {

struct PpoSampleIn {
  int action;
  float logp_old;
  float adv[DRL_MAX_REWARD_STREAMS];
  float v_old[DRL_MAX_REWARD_STREAMS];
  float vtarg[DRL_MAX_REWARD_STREAMS];
};

struct PpoSums {   // raw per-sample contributions (before the 1/nb of the means)
  float surr, ent, clip, vf;
};

__device__ __forceinline__ float crit_val(float d, int kind) {
  if (kind == 0) return d * d;
  const float a = fabsf(d);                      // SmoothL1Loss, beta = 1
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float crit_grad(float d, int kind) {
  if (kind == 0) return 2.f * d;
  return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f);
}

// logit[0..A) in, dlogit[0..A) out (if want_grad).  v_new[k], dv[k] per reward stream.
// AX / KX: action and stream counts when the launcher knows them at compile time (0 = read them from hp): the
// predicated iterations j >= A of the unrolled loops then disappear (they still take issue slots otherwise).
template <int MAXA, int AX = 0, int KX = 0>
__device__ __forceinline__ void ppo_sample(const float (&logit)[MAXA], const float* v_new,
                                           const PpoSampleIn& in, const drl_ppo_hyper_t& hp,
                                           float inv_nb, bool want_grad, float (&dlogit)[MAXA],
                                           float* dv, PpoSums& out, float& logp_out,
                                           float& ent_out) {
  const int A = AX ? AX : hp.num_actions;
  const int K = KX ? KX : hp.num_streams;
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < MAXA; ++j) if (j < A) m = fmaxf(m, logit[j]);
  float se = 0.f;
#pragma unroll
  for (int j = 0; j < MAXA; ++j) if (j < A) se += expf(logit[j] - m);
  const float lse = m + logf(se);
  float H = 0.f, logp = 0.f;
  float p[MAXA];
#pragma unroll
  for (int j = 0; j < MAXA; ++j) {
    if (j < A) {
      const float n = logit[j] - lse;
      p[j] = expf(n);
      H -= p[j] * fmaxf(n, -3.402823466e38f);
      if (j == in.action) logp = n;
    } else {
      p[j] = 0.f;
    }
  }
  logp_out = logp;
  ent_out = H;
  const float ratio = expf(logp - in.logp_old);
  const float lo = 1.f - hp.clip_param, hi = 1.f + hp.clip_param;
  const float clipped = fminf(fmaxf(ratio, lo), hi);
  float aw = 0.f;
#pragma unroll
  for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) if (k < K) aw += hp.reward_weights[k] * in.adv[k];
  const float s1 = aw * ratio, s2 = aw * clipped;
  out.surr = fminf(s1, s2);
  out.clip = s1 > s2 ? 1.f : 0.f;
  out.ent = H;
  float vf = 0.f;
#pragma unroll
  for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) {
    if (k >= K) break;
    const float coef = hp.vf_simple_weighting ? 1.f : hp.reward_weights[k] * hp.reward_weights[k];
    const float d1 = v_new[k] - in.vtarg[k];
    const float l1 = crit_val(d1, hp.vf_loss_kind);
    float l = l1, g = crit_grad(d1, hp.vf_loss_kind);
    if (hp.vf_loss_clipping) {
      const float dd = v_new[k] - in.v_old[k];
      const float cl = fminf(fmaxf(dd, -hp.clip_param), hp.clip_param);
      const float d2 = (in.v_old[k] + cl) - in.vtarg[k];
      const float l2 = crit_val(d2, hp.vf_loss_kind);
      const bool inr = dd >= -hp.clip_param && dd <= hp.clip_param;
      const float g2 = inr ? crit_grad(d2, hp.vf_loss_kind) : 0.f;
      l = fmaxf(l1, l2);
      g = l1 > l2 ? g : (l1 < l2 ? g2 : 0.5f * (g + g2));
    }
    vf += coef * l;
    if (want_grad) dv[k] = hp.vf_coef * coef * inv_nb * g;
  }
  out.vf = vf;
  if (want_grad) {
    const bool inr = ratio >= lo && ratio <= hi;
    const float g_ratio = (inr || s1 < s2) ? aw : 0.f;
    const float dlogp = -inv_nb * g_ratio * ratio;
    const float dH = -hp.ent_coef * inv_nb;
#pragma unroll
    for (int j = 0; j < MAXA; ++j) {
      if (j < A) {
        const float n = logit[j] - lse;
        dlogit[j] = dlogp * ((j == in.action ? 1.f : 0.f) - p[j]) - dH * p[j] * (fmaxf(n, -3.402823466e38f) + H);
      }
    }
  }
}

// Adds one block's partial sums into losses_out[5] = policy_loss, value_loss, composite_loss,
// meanent, clipfrac (ppo.py:195-207).  All five are linear in the raw sums.
__device__ __forceinline__ void ppo_accumulate_losses(const PpoSums& s, const drl_ppo_hyper_t& hp,
                                                      float inv_nb, float* losses_out) {
  const float policy = -(s.surr + hp.ent_coef * s.ent) * inv_nb;
  const float value = s.vf * inv_nb;
  atomicAdd(losses_out + 0, policy);
  atomicAdd(losses_out + 1, value);
  atomicAdd(losses_out + 2, policy + hp.vf_coef * value);
  atomicAdd(losses_out + 3, s.ent * inv_nb);
  atomicAdd(losses_out + 4, s.clip * inv_nb);
}

// Block-wide reduction of PpoSums (every thread contributes); thread 0 gets the result.
__device__ __forceinline__ PpoSums block_reduce_sums(PpoSums v, float* smem /* >= 4*32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v.surr = warp_sum(v.surr); v.ent = warp_sum(v.ent); v.clip = warp_sum(v.clip); v.vf = warp_sum(v.vf);
  if (lane == 0) {
    smem[warp] = v.surr; smem[32 + warp] = v.ent; smem[64 + warp] = v.clip; smem[96 + warp] = v.vf;
  }
  __syncthreads();
  PpoSums r{0.f, 0.f, 0.f, 0.f};
  if (warp == 0) {
    r.surr = warp_sum(lane < nw ? smem[lane] : 0.f);
    r.ent = warp_sum(lane < nw ? smem[32 + lane] : 0.f);
    r.clip = warp_sum(lane < nw ? smem[64 + lane] : 0.f);
    r.vf = warp_sum(lane < nw ? smem[96 + lane] : 0.f);
  }
  return r;
}

__device__ __forceinline__ void load_sample(const drl_ppo_batch_t& b, int K, int64_t r, PpoSampleIn& in) {
  in.action = (int)b.actions[r];
  in.logp_old = b.logp_old[r];
#pragma unroll
  for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) {
    if (k < K) {
      in.adv[k] = b.adv[k][r];
      in.v_old[k] = b.v_old[k] ? b.v_old[k][r] : 0.f;
      in.vtarg[k] = b.vtarg[k][r];
    }
  }
}

}  // namespace drl
