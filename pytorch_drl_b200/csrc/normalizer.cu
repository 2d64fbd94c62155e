// Observation normaliser: replaces Normalizer.forward / update
// (drl/envs/wrappers/stateful/normalize.py:129-162, 105-127).  Element-wise, HBM-bound.
#include "common.cuh"

namespace drl {

__global__ void __launch_bounds__(256) normalizer_apply_kernel(
    const float* __restrict__ x, const float* __restrict__ m1, const float* __restrict__ m2,
    int64_t total, int64_t d, float eps, float lo, float hi, float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = i % d;
    const float mu = __ldg(m1 + j);
    const float var = __fsub_rn(__ldg(m2 + j), __fmul_rn(mu, mu));
    float v = __fmul_rn(__fsub_rn(x[i], mu), rsqrtf(__fadd_rn(var, eps)));
    v = fminf(fmaxf(v, lo), hi);
    y[i] = v;
  }
}

// Running non-central moments, items applied in order (normalize.py:115-123):
//   steps += 1; m *= (steps-1)/steps; m += (1/steps) * item   (all fp32, no contraction)
__global__ void __launch_bounds__(256) normalizer_update_kernel(
    float* __restrict__ m1, float* __restrict__ m2, const float* __restrict__ x, int64_t count,
    int64_t d, float steps_before) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < d;
       j += (int64_t)gridDim.x * blockDim.x) {
    float a1 = m1[j], a2 = m2[j];
    float steps = steps_before;
    for (int64_t i = 0; i < count; ++i) {
      steps = __fadd_rn(steps, 1.f);
      const float keep = __fdiv_rn(__fsub_rn(steps, 1.f), steps);
      const float w = __fdiv_rn(1.f, steps);
      const float v = x[i * d + j];
      a1 = __fadd_rn(__fmul_rn(a1, keep), __fmul_rn(w, v));
      a2 = __fadd_rn(__fmul_rn(a2, keep), __fmul_rn(w, __fmul_rn(v, v)));
    }
    m1[j] = a1;
    m2[j] = a2;
  }
}

}  // namespace drl

using namespace drl;

extern "C" int drl_normalizer_apply_f32(const float* x, const float* m1, const float* m2, int64_t n,
                                        int64_t d, float eps, float clip_lo, float clip_hi,
                                        float* y, void* stream) {
  if (!x || !m1 || !m2 || !y) return DRL_ERR_BAD_ARG;
  if (n < 0 || d < 1) return DRL_ERR_BAD_SHAPE;
  if (n == 0) return DRL_OK;
  const int64_t total = n * d;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  normalizer_apply_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, m1, m2, total, d, eps,
                                                                          clip_lo, clip_hi, y);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_normalizer_update_f32(float* m1, float* m2, const float* x, int64_t count,
                                         int64_t d, double steps_before, void* stream) {
  if (!m1 || !m2 || (count > 0 && !x)) return DRL_ERR_BAD_ARG;
  if (count < 0 || d < 1 || steps_before < 0) return DRL_ERR_BAD_SHAPE;
  if (count == 0) return DRL_OK;
  int64_t blocks = (d + 255) / 256;
  normalizer_update_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(m1, m2, x, count, d,
                                                                           (float)steps_before);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
