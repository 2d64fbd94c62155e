// Flat fused Adam: replaces torch.optim.Adam.step (configured by drl/utils/optimization.py:9-35;
// models_dir/atari_ppo/config.yaml:76-78) on one contiguous fp32 buffer.  One float4 pass,
// 28 B/param of HBM traffic (read p,g,m,v; write p,m,v).  grad_scale folds the 1/world of the
// DDP mean (drl/utils/launch.py:310) into the update.
#include "common.cuh"

namespace drl {

struct AdamCoef { float lr_bc1, inv_sqrt_bc2, beta1, beta2, omb1, omb2, eps, gscale; };

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamCoef& c) {
  g *= c.gscale;
  m = m + (g - m) * c.omb1;                 // exp_avg.lerp_(grad, 1-beta1)
  v = v * c.beta2 + c.omb2 * g * g;         // mul_(beta2).addcmul_(g, g, 1-beta2)
  const float denom = sqrtf(v) * c.inv_sqrt_bc2 + c.eps;
  p = p - c.lr_bc1 * (m / denom);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   int64_t n, AdamCoef c, int vec_ok) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t n4 = vec_ok ? n / 4 : 0;
  for (int64_t i = tid; i < n4; i += stride) {
    float4 p4 = reinterpret_cast<float4*>(p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    adam_one(p4.x, g4.x, m4.x, v4.x, c);
    adam_one(p4.y, g4.y, m4.y, v4.y, c);
    adam_one(p4.z, g4.z, m4.z, v4.z, c);
    adam_one(p4.w, g4.w, m4.w, v4.w, c);
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], c);
}

int adam_launch(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                double lr, double beta1, double beta2, double eps, int64_t step, float grad_scale,
                cudaStream_t st) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) return DRL_ERR_BAD_ARG;
  if (n < 0 || step < 1) return DRL_ERR_BAD_SHAPE;
  if (n == 0) return DRL_OK;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  // 1-beta is formed in double like torch's python scalars (value=1-beta2), then rounded once
  AdamCoef c{(float)(lr / bc1), (float)(1.0 / sqrt(bc2)), (float)beta1, (float)beta2,
             (float)(1.0 - beta1), (float)(1.0 - beta2), (float)eps, grad_scale};
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const int vec_ok = al(params) && al(grads) && al(exp_avg) && al(exp_avg_sq);
  const int threads = 256;
  int64_t want = (n / 4 + threads - 1) / threads + 1;
  int blocks = (int)(want < (int64_t)sm_count() * 8 ? want : (int64_t)sm_count() * 8);
  adam_kernel<<<blocks, threads, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, c, vec_ok);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

}  // namespace drl

extern "C" int drl_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                             int64_t n, double lr, double beta1, double beta2, double eps,
                             int64_t step, float grad_scale, void* stream) {
  return drl::adam_launch(params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step,
                          grad_scale, drl::as_stream(stream));
}
