// GAE(lambda) + value targets + per-env standardisation: one warp per env (trajectory).
//
// Replaces GAE.call (drl/algos/common/credit_assignment.py:195-214), the tail of
// ActorCriticAlgo.credit_assignment (drl/algos/abstract.py:427-439) and standardize
// (drl/utils/stats.py:4-17).
//
// The recurrence A_t = delta_t + c_t * A_{t+1} is affine in A_{t+1}, so a lane folds its own
// contiguous chunk of timesteps into one map (a, b), the 32 maps are combined by a 5-step
// suffix scan with shuffles, and every lane then replays its chunk from the incoming value.
// HBM traffic is the algorithmic 20 B per (env, step): r, V, d read once (float4 per lane when
// T is 128 or 256), A and vtarg written once.
#include "common.cuh"

namespace drl {

struct Affine { float a, b; };   // x -> a + b * x

__device__ __forceinline__ Affine compose(Affine f, Affine g) {  // f o g
  return Affine{fmaf(f.b, g.a, f.a), f.b * g.b};
}

// Suffix scan over lanes: returns X_l = (F_{l+1} o ... o F_31)(0).
__device__ __forceinline__ float suffix_incoming(Affine f, int lane) {
  Affine s = f;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    Affine o;
    o.a = __shfl_down_sync(0xffffffffu, s.a, off);
    o.b = __shfl_down_sync(0xffffffffu, s.b, off);
    if (lane + off < 32) s = compose(s, o);
  }
  float x = __shfl_down_sync(0xffffffffu, s.a, 1);   // S_{l+1}(0) = a of the inclusive suffix
  return lane == 31 ? 0.f : x;
}

__device__ __forceinline__ void step_coeffs(float r, float v, float vn, float d, float gamma,
                                            float gl, float lambda, int use_dones,
                                            float& delta, float& c) {
  // same association as the reference: (-V_t + r_t) + (nt*gamma)*V_{t+1}; ((nt*gamma)*lambda)
  float cv, ca;
  if (use_dones) {
    float nt = 1.f - d;
    cv = nt * gamma;
    ca = cv * lambda;
  } else {
    cv = gamma;
    ca = gl;
  }
  delta = __fadd_rn(__fadd_rn(-v, r), __fmul_rn(cv, vn));
  c = ca;
}

// C = timesteps per lane (T == 32*C), vectorised loads of C floats per lane.
template <int C>
__global__ void __launch_bounds__(128) gae_warp_vec_kernel(
    const float* __restrict__ rewards, const float* __restrict__ vpreds,
    const float* __restrict__ dones, int n_envs, int ld_r, int ld_v, int ld_o,
    float gamma, float lambda, float gl, int use_dones, int standardize,
    float* __restrict__ adv_out, float* __restrict__ vtarg_out) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= n_envs) return;
  constexpr int T = 32 * C;
  const float* r_row = rewards + (size_t)env * ld_r;
  const float* v_row = vpreds + (size_t)env * ld_v;
  const float* d_row = dones + (size_t)env * ld_r;
  float r[C], v[C + 1], d[C];
#pragma unroll
  for (int q = 0; q < C / 4; ++q) {
    float4 r4 = __ldg(reinterpret_cast<const float4*>(r_row) + lane * (C / 4) + q);
    float4 v4 = __ldg(reinterpret_cast<const float4*>(v_row) + lane * (C / 4) + q);
    float4 d4 = use_dones ? __ldg(reinterpret_cast<const float4*>(d_row) + lane * (C / 4) + q)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    r[4 * q] = r4.x; r[4 * q + 1] = r4.y; r[4 * q + 2] = r4.z; r[4 * q + 3] = r4.w;
    v[4 * q] = v4.x; v[4 * q + 1] = v4.y; v[4 * q + 2] = v4.z; v[4 * q + 3] = v4.w;
    d[4 * q] = d4.x; d[4 * q + 1] = d4.y; d[4 * q + 2] = d4.z; d[4 * q + 3] = d4.w;
  }
  // V at the first step of the next lane's chunk; lane 31 reads the bootstrap value V_T.
  float vnext = __shfl_down_sync(0xffffffffu, v[0], 1);
  if (lane == 31) vnext = __ldg(v_row + T);
  v[C] = vnext;

  float delta[C], c[C];
  Affine f{0.f, 1.f};
#pragma unroll
  for (int i = C - 1; i >= 0; --i) {
    step_coeffs(r[i], v[i], v[i + 1], d[i], gamma, gl, lambda, use_dones, delta[i], c[i]);
    f = Affine{fmaf(c[i], f.a, delta[i]), c[i] * f.b};
  }
  float x = suffix_incoming(f, lane);
  float adv[C];
#pragma unroll
  for (int i = C - 1; i >= 0; --i) {
    x = __fadd_rn(delta[i], __fmul_rn(c[i], x));
    adv[i] = x;
  }
  float* vt_row = vtarg_out + (size_t)env * ld_o;
  float* a_row = adv_out + (size_t)env * ld_o;
#pragma unroll
  for (int q = 0; q < C / 4; ++q) {
    float4 o = make_float4(adv[4 * q] + v[4 * q], adv[4 * q + 1] + v[4 * q + 1],
                           adv[4 * q + 2] + v[4 * q + 2], adv[4 * q + 3] + v[4 * q + 3]);
    reinterpret_cast<float4*>(vt_row)[lane * (C / 4) + q] = o;
  }
  if (standardize) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) s += adv[i];
    const float mean = warp_sum(s) / (float)T;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) { float e = adv[i] - mean; ss = fmaf(e, e, ss); }
    const float stdv = sqrtf(warp_sum(ss) / (float)(T - 1));   // unbiased, no eps (stats.py:16)
#pragma unroll
    for (int i = 0; i < C; ++i) adv[i] = (adv[i] - mean) / stdv;
  }
#pragma unroll
  for (int q = 0; q < C / 4; ++q) {
    reinterpret_cast<float4*>(a_row)[lane * (C / 4) + q] =
        make_float4(adv[4 * q], adv[4 * q + 1], adv[4 * q + 2], adv[4 * q + 3]);
  }
}

// Any T >= 1: lane l owns timesteps [l*c, min(T,(l+1)*c)), c = ceil(T/32); three streaming passes
// (fold, replay+write, standardise) — the second and third hit L1/L2.
__global__ void __launch_bounds__(128) gae_warp_generic_kernel(
    const float* __restrict__ rewards, const float* __restrict__ vpreds,
    const float* __restrict__ dones, int n_envs, int T, int ld_r, int ld_v, int ld_o,
    float gamma, float lambda, float gl, int use_dones, int standardize,
    float* __restrict__ adv_out, float* __restrict__ vtarg_out) {
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= n_envs) return;
  const int c = (T + 31) / 32;
  const int s = min(T, lane * c), e = min(T, s + c);
  const float* r_row = rewards + (size_t)env * ld_r;
  const float* v_row = vpreds + (size_t)env * ld_v;
  const float* d_row = dones + (size_t)env * ld_r;
  float* a_row = adv_out + (size_t)env * ld_o;
  float* vt_row = vtarg_out + (size_t)env * ld_o;
  Affine f{0.f, 1.f};
  for (int t = e - 1; t >= s; --t) {
    float delta, cc;
    step_coeffs(r_row[t], v_row[t], v_row[t + 1], use_dones ? d_row[t] : 0.f, gamma, gl, lambda,
                use_dones, delta, cc);
    f = Affine{fmaf(cc, f.a, delta), cc * f.b};
  }
  float x = suffix_incoming(f, lane);
  float sum = 0.f;
  for (int t = e - 1; t >= s; --t) {
    float delta, cc;
    const float v = v_row[t];
    step_coeffs(r_row[t], v, v_row[t + 1], use_dones ? d_row[t] : 0.f, gamma, gl, lambda,
                use_dones, delta, cc);
    x = __fadd_rn(delta, __fmul_rn(cc, x));
    a_row[t] = x;
    vt_row[t] = x + v;
    sum += x;
  }
  if (standardize) {
    const float mean = warp_sum(sum) / (float)T;
    float ss = 0.f;
    for (int t = s; t < e; ++t) { float q = a_row[t] - mean; ss = fmaf(q, q, ss); }
    const float stdv = sqrtf(warp_sum(ss) / (float)(T - 1));
    for (int t = s; t < e; ++t) a_row[t] = (a_row[t] - mean) / stdv;
  }
}

__global__ void __launch_bounds__(128) standardize_warp_kernel(const float* __restrict__ x,
                                                              int n_rows, int T, int ld,
                                                              float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const float* xr = x + (size_t)row * ld;
  float* yr = y + (size_t)row * ld;
  float s = 0.f;
  for (int t = lane; t < T; t += 32) s += xr[t];
  const float mean = warp_sum(s) / (float)T;
  float ss = 0.f;
  for (int t = lane; t < T; t += 32) { float q = xr[t] - mean; ss = fmaf(q, q, ss); }
  const float stdv = sqrtf(warp_sum(ss) / (float)(T - 1));
  for (int t = lane; t < T; t += 32) yr[t] = (xr[t] - mean) / stdv;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace drl

using namespace drl;

extern "C" int drl_gae_f32(const float* rewards, const float* vpreds, const float* dones,
                           int n_envs, int T, int ld_r, int ld_v, int ld_o, float gamma,
                           float lambda, int use_dones, int standardize, float* adv_out,
                           float* vtarg_out, void* stream) {
  if (!rewards || !vpreds || !adv_out || !vtarg_out || (use_dones && !dones)) return DRL_ERR_BAD_ARG;
  if (n_envs < 0 || T < 1 || ld_r < T || ld_v < T + 1 || ld_o < T) return DRL_ERR_BAD_SHAPE;
  if (n_envs == 0) return DRL_OK;
  if (!dones) dones = rewards;   // never dereferenced when !use_dones
  const float gl = (float)((double)gamma * (double)lambda);
  const int warps = 4;
  dim3 grid((n_envs + warps - 1) / warps), block(32 * warps);
  const bool vec_ok = (ld_r % 4 == 0) && (ld_v % 4 == 0) && (ld_o % 4 == 0) && aligned16(rewards) &&
                      aligned16(vpreds) && aligned16(dones) && aligned16(adv_out) &&
                      aligned16(vtarg_out);
  cudaStream_t st = as_stream(stream);
  if (vec_ok && T == 128) {
    gae_warp_vec_kernel<4><<<grid, block, 0, st>>>(rewards, vpreds, dones, n_envs, ld_r, ld_v, ld_o,
                                                   gamma, lambda, gl, use_dones, standardize,
                                                   adv_out, vtarg_out);
  } else if (vec_ok && T == 256) {
    gae_warp_vec_kernel<8><<<grid, block, 0, st>>>(rewards, vpreds, dones, n_envs, ld_r, ld_v, ld_o,
                                                   gamma, lambda, gl, use_dones, standardize,
                                                   adv_out, vtarg_out);
  } else {
    gae_warp_generic_kernel<<<grid, block, 0, st>>>(rewards, vpreds, dones, n_envs, T, ld_r, ld_v,
                                                    ld_o, gamma, lambda, gl, use_dones,
                                                    standardize, adv_out, vtarg_out);
  }
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_standardize_f32(const float* x, int n_rows, int T, int ld, float* y,
                                   void* stream) {
  if (!x || !y) return DRL_ERR_BAD_ARG;
  if (n_rows < 0 || T < 2 || ld < T) return DRL_ERR_BAD_SHAPE;
  if (n_rows == 0) return DRL_OK;
  const int warps = 4;
  standardize_warp_kernel<<<(n_rows + warps - 1) / warps, 32 * warps, 0, as_stream(stream)>>>(
      x, n_rows, T, ld, y);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
