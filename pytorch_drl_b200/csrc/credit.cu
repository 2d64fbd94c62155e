// n-step credit assignment — replaces NStepAdvantageEstimator.call and
// SimpleDiscreteBellmanOptimalityOperator.call (drl/algos/common/credit_assignment.py:225-249, 276-305).
// Every (env, t) is independent: one thread each walks its n = extra_steps + 1 rewards backwards from the
// bootstrap value, R = r + nonterminal * gamma * R in float32 with the reference's operation order (separate
// multiply and add, no FMA contraction), so results are bit-identical to the reference.
#include "common.cuh"

namespace drl {

__device__ __forceinline__ float nstep_backup(float R, const float* __restrict__ r, const float* __restrict__ d, int n,
                                               float gamma, int use_dones) {
  for (int s = n - 1; s >= 0; --s) {
    const float ntg = use_dones ? __fmul_rn(__fsub_rn(1.f, d[s]), gamma) : gamma;
    R = __fadd_rn(r[s], __fmul_rn(ntg, R));
  }
  return R;
}

__global__ void __launch_bounds__(256) nstep_adv_kernel(const float* __restrict__ rewards, const float* __restrict__ vpreds,
                                                         const float* __restrict__ dones, int N, int T, int n, int64_t ld_r,
                                                         int64_t ld_v, int64_t ld_o, float gamma, int use_dones,
                                                         float* __restrict__ adv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * T) return;
  const int e = (int)(i / T), t = (int)(i - (int64_t)e * T);
  const float R = nstep_backup(vpreds[e * ld_v + t + n], rewards + e * ld_r + t, dones + e * ld_r + t, n, gamma, use_dones);
  adv[e * ld_o + t] = __fsub_rn(R, vpreds[e * ld_v + t]);
}

__global__ void __launch_bounds__(256) bellman_kernel(const float* __restrict__ rewards, const float* __restrict__ q,
                                                       const float* __restrict__ q_tgt, const float* __restrict__ dones, int N,
                                                       int T, int n, int A, int64_t ld_r, int64_t ld_q, int64_t ld_o,
                                                       float gamma, int use_dones, int double_q, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)N * T) return;
  const int e = (int)(i / T), t = (int)(i - (int64_t)e * T);
  const float* qt = q_tgt + (e * ld_q + t + n) * A;
  const float* qs = double_q ? q + (e * ld_q + t + n) * A : qt;
  int a = 0;
  float best = qs[0];
  for (int j = 1; j < A; ++j) {
    const float v = qs[j];
    if (v > best) { best = v; a = j; }                 // first maximum, like torch.argmax
  }
  out[e * ld_o + t] = nstep_backup(qt[a], rewards + e * ld_r + t, dones + e * ld_r + t, n, gamma, use_dones);
}

}  // namespace drl

using namespace drl;

extern "C" int drl_nstep_advantages_f32(const float* rewards, const float* vpreds, const float* dones, int n_envs, int T,
                                        int extra_steps, int64_t ld_r, int64_t ld_v, int64_t ld_out, float gamma,
                                        int use_dones, float* adv_out, void* stream) {
  if (!rewards || !vpreds || !dones || !adv_out) return DRL_ERR_BAD_ARG;
  const int n = extra_steps + 1;
  if (n_envs < 1 || T < 1 || extra_steps < 0 || ld_r < T + n - 1 || ld_v < T + n || ld_out < T) return DRL_ERR_BAD_SHAPE;
  const int64_t total = (int64_t)n_envs * T;
  nstep_adv_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(rewards, vpreds, dones, n_envs, T, n, ld_r,
                                                                                    ld_v, ld_out, gamma, use_dones, adv_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_bellman_backups_f32(const float* rewards, const float* qpreds, const float* tgt_qpreds, const float* dones,
                                       int n_envs, int T, int extra_steps, int num_actions, int64_t ld_r, int64_t ld_q,
                                       int64_t ld_out, float gamma, int use_dones, int double_q, float* out, void* stream) {
  if (!rewards || !tgt_qpreds || !dones || !out || (double_q && !qpreds)) return DRL_ERR_BAD_ARG;
  const int n = extra_steps + 1;
  if (n_envs < 1 || T < 1 || extra_steps < 0 || num_actions < 1 || ld_r < T + n - 1 || ld_q < T + n || ld_out < T)
    return DRL_ERR_BAD_SHAPE;
  const int64_t total = (int64_t)n_envs * T;
  bellman_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(rewards, qpreds, tgt_qpreds, dones, n_envs, T, n,
                                                                                  num_actions, ld_r, ld_q, ld_out, gamma,
                                                                                  use_dones, double_q, out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
