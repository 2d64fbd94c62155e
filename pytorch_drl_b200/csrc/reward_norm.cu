// Reward normaliser for N environments at once — replaces ReturnAcc.update/_unload/_update_from_returns/forward
// and the per-step part of NormalizeRewardWrapper.step (drl/envs/wrappers/stateful/normalize_reward.py:20-100,
// 171-194).  State per environment: a window of pending (reward, done) pairs stored time-major
// ([2L][N] so that the N threads of a step touch consecutive addresses), its fill count, and the running
// moments (steps, moment1, moment2) of its un-synchronised accumulator.
//
// One thread per environment walks its T new steps in order (the window update is a strictly sequential
// backward scan — the work is ~20 B of traffic per step; this is integer/byte work, not a GEMM):
//   append (r_t, d_t); when 2L pairs are pending: returns_t = r_t + mask_t * gamma * returns_{t+1} backwards from
//   a zero bootstrap over the whole window (float32, gamma applied in float32 like the reference's
//   python-scalar x float32-tensor product), mean and mean-square of the FIRST L returns are folded into the
//   moments with weight L, and those L pairs are dropped.
// The normalised reward uses the SYNCHRONISED moments (one set per rank, refreshed by learn()):
//   0 if steps == 0, else clip(r * rsqrt(moment2 - moment1^2 + eps), lo, hi)   (shift = False, :97).
#include "common.cuh"

namespace drl {

__global__ void __launch_bounds__(128) return_acc_update_kernel(const float* __restrict__ rewards, int64_t r_ld,
                                                                 const float* __restrict__ dones, int64_t d_ld, int N,
                                                                 int T, float gamma, int use_dones, int L,
                                                                 float* __restrict__ win_r, float* __restrict__ win_d,
                                                                 int* __restrict__ count, float* __restrict__ steps,
                                                                 float* __restrict__ m1, float* __restrict__ m2) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N) return;
  int cnt = count[e];
  float st = steps[e], a1 = m1[e], a2 = m2[e];
  const int W = 2 * L;
  for (int t = 0; t < T; ++t) {
    win_r[(int64_t)cnt * N + e] = rewards[(int64_t)e * r_ld + t];
    win_d[(int64_t)cnt * N + e] = dones[(int64_t)e * d_ld + t];
    if (++cnt < W) continue;
    // ---- unload (normalize_reward.py:60-68) ----
    float ret = 0.f;
    double sum = 0.0, sumsq = 0.0;
    for (int k = W - 1; k >= 0; --k) {
      const float r = win_r[(int64_t)k * N + e];
      const bool done = use_dones && win_d[(int64_t)k * N + e] != 0.f;
      ret = done ? r : __fadd_rn(r, __fmul_rn(gamma, ret));
      if (k < L) { sum += (double)ret; sumsq += (double)__fmul_rn(ret, ret); }
    }
    for (int k = 0; k < L; ++k) {                       // keep the newer half
      win_r[(int64_t)k * N + e] = win_r[(int64_t)(k + L) * N + e];
      win_d[(int64_t)k * N + e] = win_d[(int64_t)(k + L) * N + e];
    }
    cnt = L;
    // ---- moments (normalize_reward.py:70-88), float32 in the reference's order ----
    const float Lf = (float)L;
    const float ns = __fadd_rn(st, Lf);
    const float keep = __fdiv_rn(__fsub_rn(ns, Lf), ns), w = __fdiv_rn(Lf, ns);
    const float mean = (float)(sum / (double)L), msq = (float)(sumsq / (double)L);
    a1 = __fadd_rn(__fmul_rn(a1, keep), __fmul_rn(w, mean));
    a2 = __fadd_rn(__fmul_rn(a2, keep), __fmul_rn(w, msq));
    st = ns;
  }
  count[e] = cnt; steps[e] = st; m1[e] = a1; m2[e] = a2;
}

__global__ void __launch_bounds__(256) reward_norm_apply_kernel(const float* __restrict__ x, int64_t n,
                                                                 const float* __restrict__ synced /* steps, m1, m2 */,
                                                                 float eps, float lo, float hi, float* __restrict__ y) {
  const float st = synced[0], a1 = synced[1], a2 = synced[2];
  const float inv = rsqrtf(__fadd_rn(__fsub_rn(a2, __fmul_rn(a1, a1)), eps));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = st == 0.f ? 0.f : fminf(fmaxf(__fmul_rn(x[i], inv), lo), hi);
}

}  // namespace drl

using namespace drl;

extern "C" int drl_return_acc_update_f32(const float* rewards, int64_t r_ld, const float* dones, int64_t d_ld, int n_envs,
                                         int T, float gamma, int use_dones, int trace_len, float* win_r, float* win_d,
                                         int* count, float* steps, float* m1, float* m2, void* stream) {
  if (!rewards || !dones || !win_r || !win_d || !count || !steps || !m1 || !m2) return DRL_ERR_BAD_ARG;
  if (n_envs < 1 || T < 0 || trace_len < 1 || r_ld < T || d_ld < T) return DRL_ERR_BAD_SHAPE;
  if (T == 0) return DRL_OK;
  return_acc_update_kernel<<<(n_envs + 127) / 128, 128, 0, as_stream(stream)>>>(rewards, r_ld, dones, d_ld, n_envs, T, gamma,
                                                                                use_dones, trace_len, win_r, win_d, count,
                                                                                steps, m1, m2);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_reward_norm_apply_f32(const float* x, int64_t n, const float* synced_moments, float eps, float clip_lo,
                                         float clip_hi, float* y, void* stream) {
  if (!x || !y || !synced_moments) return DRL_ERR_BAD_ARG;
  if (n < 0) return DRL_ERR_BAD_SHAPE;
  if (n == 0) return DRL_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  reward_norm_apply_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, n, synced_moments, eps, clip_lo, clip_hi, y);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
