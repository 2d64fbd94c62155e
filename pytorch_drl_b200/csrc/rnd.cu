// RND predictor path: replaces the arithmetic of RandomNetworkDistillationWrapper.learn / .step
// (drl/envs/wrappers/stateful/intrinsic/random_network_distillation.py:209-235, 186-201) with the
// _TeacherNetwork / _StudentNetwork of the same file (:22-87).
//
// Two engines behind one handle: DRL_PRECISION_FP32 = the fp32 CUDA-core implicit GEMM (simt_conv.cu), the strict-parity
// mode (the uint8 frame's last channel is read directly; x*(1/255), the observation normaliser (normalize.py:149-162) and
// the +-5 clip are fused into the first conv's gather; the predictor MSE forward/backward is one warp-per-sample kernel);
// DRL_PRECISION_BF16 = the tcgen05 engine of umma_rnd.cu (throughput mode).
#include <new>
#include "rnd.cuh"
#include "simt_conv.cuh"
#include "net.cuh"

namespace drl {

// per-sample error err[i] = mean_d (y - yh)^2; loss += err[i]/nb; dyh = 2 (yh - y) / (D nb)
__global__ void __launch_bounds__(256) rnd_mse_kernel(const float* __restrict__ y, const float* __restrict__ yh,
                                                      int nb, float* __restrict__ err, float* __restrict__ loss,
                                                      float* __restrict__ dyh) {
  constexpr int D = 512;
  const int lane = threadIdx.x & 31, wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  float lacc = 0.f;
  for (int i = wid; i < nb; i += nw) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < D / 32; ++q) {
      const size_t e = (size_t)i * D + lane + 32 * q;
      const float d = y[e] - yh[e];
      s = fmaf(d, d, s);
      if (dyh) dyh[e] = -2.f * d / ((float)D * (float)nb);
    }
    s = warp_sum(s) / (float)D;
    if (lane == 0) {
      if (err) err[i] = s;
      lacc += s;
    }
  }
  if (loss && lane == 0 && lacc != 0.f) atomicAdd(loss, lacc / (float)nb);
}

static const ConvGeom kTrunk[3] = {ConvGeom{84, 84, 1, 20, 20, 16, 8, 8, 4}, ConvGeom{20, 20, 16, 9, 9, 32, 4, 4, 2},
                                   ConvGeom{9, 9, 32, 6, 6, 32, 4, 4, 1}};
static const ConvGeom kTeacherFc = ConvGeom{6, 6, 32, 1, 1, 512, 6, 6, 1};
static const ConvGeom kStudentFc[3] = {ConvGeom{6, 6, 32, 1, 1, 256, 6, 6, 1}, ConvGeom{1, 1, 256, 1, 1, 256, 1, 1, 1},
                                       ConvGeom{1, 1, 256, 1, 1, 512, 1, 1, 1}};

}  // namespace drl

using namespace drl;

static void rnd_offsets(drl_rnd* r) {
  const int64_t trunk[6] = {16 * 64, 16, 32 * 16 * 16, 32, 32 * 32 * 16, 32};
  int64_t o = 0;
  for (int i = 0; i < 6; ++i) { r->toff[i] = o; r->soff[i] = o; o += trunk[i]; }
  r->toff[6] = o; r->toff[7] = o + 512 * 1152; r->toff[8] = o + 512 * 1152 + 512;
  const int64_t st[6] = {256 * 1152, 256, 256 * 256, 256, 512 * 256, 512};
  int64_t so = o;
  for (int i = 0; i < 6; ++i) { r->soff[6 + i] = so; so += st[i]; }
  r->soff[12] = so;
}

extern "C" int64_t drl_rnd_param_count(int student, int widening) {
  if (widening != 1) return -1;
  const int64_t trunk = 16 * 64 + 16 + 32 * 16 * 16 + 32 + 32 * 32 * 16 + 32;
  return student ? trunk + 256 * 1152 + 256 + 256 * 256 + 256 + 512 * 256 + 512 : trunk + 512 * 1152 + 512;
}

extern "C" int drl_rnd_create(drl_rnd_t** out, int device, int max_batch, int widening, int precision) {
  if (!out) return DRL_ERR_BAD_ARG;
  *out = nullptr;
  if (max_batch < 1) return DRL_ERR_BAD_SHAPE;
  if (widening != 1) return DRL_ERR_UNSUPPORTED;
  if (precision != DRL_PRECISION_FP32 && precision != DRL_PRECISION_BF16) return DRL_ERR_BAD_ARG;
  DRL_TRY(drl_check_device(device));
  DRL_CUDA(cudaSetDevice(device));
  drl_rnd* r = new (std::nothrow) drl_rnd();
  if (!r) return DRL_ERR_CUDA;
  r->device = device; r->max_batch = max_batch; r->precision = precision;
  rnd_offsets(r);
  if (precision == DRL_PRECISION_BF16) {
    const int s = rnd_bf16_alloc(r);
    if (s != DRL_OK) { drl_rnd_destroy(r); return s; }
    *out = r;
    return DRL_OK;
  }
  const size_t nb = (size_t)max_batch;
  const size_t e[3] = {nb * 400 * 16, nb * 81 * 32, nb * 36 * 32};
  auto al = [](float** p, size_t n) { return cudaMalloc(p, n * sizeof(float)); };
  cudaError_t err = cudaSuccess;
  for (int i = 0; i < 3 && err == cudaSuccess; ++i) {
    err = al(&r->ta[i], e[i]);
    if (err == cudaSuccess) err = al(&r->sa[i], e[i]);
    if (err == cudaSuccess) err = al(&r->dsa[i], e[i]);
  }
  for (int i = 0; i < 2 && err == cudaSuccess; ++i) {
    err = al(&r->sh[i], nb * 256);
    if (err == cudaSuccess) err = al(&r->dsh[i], nb * 256);
  }
  if (err == cudaSuccess) err = al(&r->y, nb * 512);
  if (err == cudaSuccess) err = al(&r->yh, nb * 512);
  if (err == cudaSuccess) err = al(&r->dyh, nb * 512);
  if (err != cudaSuccess) { drl_rnd_destroy(r); return cuda_fail(err); }
  *out = r;
  return DRL_OK;
}

extern "C" int drl_rnd_destroy(drl_rnd_t* r) {
  if (!r) return DRL_OK;
  cudaSetDevice(r->device);
  for (int i = 0; i < 3; ++i) { cudaFree(r->ta[i]); cudaFree(r->sa[i]); cudaFree(r->dsa[i]); }
  for (int i = 0; i < 2; ++i) { cudaFree(r->sh[i]); cudaFree(r->dsh[i]); }
  cudaFree(r->y); cudaFree(r->yh); cudaFree(r->dyh);
  rnd_bf16_free(r);
  delete r;
  return DRL_OK;
}

extern "C" int drl_rnd_set_params(drl_rnd_t* r, const float* teacher_params, const float* student_params, void* stream) {
  if (!r || !teacher_params || !student_params) return DRL_ERR_BAD_ARG;
  r->tp = teacher_params; r->sp = student_params;
  if (r->precision == DRL_PRECISION_BF16) return rnd_bf16_pack(r, true, true, as_stream(stream));
  return DRL_OK;
}

extern "C" int drl_rnd_adam_step(drl_rnd_t* r, float* student_params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                 double lr, double beta1, double beta2, double eps, int64_t step, float grad_scale, void* stream) {
  if (!r || !r->tp) return DRL_ERR_BAD_ARG;
  {
    ProfileScope prof("rnd_adam", as_stream(stream));
    DRL_TRY(adam_launch(student_params, grads, exp_avg, exp_avg_sq, r->soff[12], lr, beta1, beta2, eps, step, grad_scale,
                        as_stream(stream)));
  }
  r->sp = student_params;
  if (r->precision == DRL_PRECISION_BF16) return rnd_bf16_pack(r, false, true, as_stream(stream));
  return DRL_OK;
}

static int rnd_forward(drl_rnd* r, const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1, const float* m2,
                       cudaStream_t st) {
  const float *tp = r->tp, *sp = r->sp;
  const int64_t *to = r->toff, *so = r->soff;
  // teacher
  DRL_TRY(simt_rnd_conv1_forward(obs, row_idx, nb, m1, m2, kTrunk[0], tp + to[0], tp + to[1], r->ta[0], st));
  DRL_TRY(simt_conv_forward<float>(r->ta[0], nullptr, nb, kTrunk[1], tp + to[2], tp + to[3], r->ta[1], st, 2));
  DRL_TRY(simt_conv_forward<float>(r->ta[1], nullptr, nb, kTrunk[2], tp + to[4], tp + to[5], r->ta[2], st, 2));
  DRL_TRY(simt_conv_forward<float>(r->ta[2], nullptr, nb, kTeacherFc, tp + to[6], tp + to[7], r->y, st, 0));
  // student
  DRL_TRY(simt_rnd_conv1_forward(obs, row_idx, nb, m1, m2, kTrunk[0], sp + so[0], sp + so[1], r->sa[0], st));
  DRL_TRY(simt_conv_forward<float>(r->sa[0], nullptr, nb, kTrunk[1], sp + so[2], sp + so[3], r->sa[1], st, 2));
  DRL_TRY(simt_conv_forward<float>(r->sa[1], nullptr, nb, kTrunk[2], sp + so[4], sp + so[5], r->sa[2], st, 2));
  DRL_TRY(simt_conv_forward<float>(r->sa[2], nullptr, nb, kStudentFc[0], sp + so[6], sp + so[7], r->sh[0], st, 1));
  DRL_TRY(simt_conv_forward<float>(r->sh[0], nullptr, nb, kStudentFc[1], sp + so[8], sp + so[9], r->sh[1], st, 1));
  DRL_TRY(simt_conv_forward<float>(r->sh[1], nullptr, nb, kStudentFc[2], sp + so[10], sp + so[11], r->yh, st, 0));
  return DRL_OK;
}

static int rnd_check(drl_rnd* r, const void* obs, int nb, const float* m1, const float* m2) {
  if (!r || !obs || !m1 || !m2 || !r->tp || !r->sp) return DRL_ERR_BAD_ARG;
  if (nb < 1 || nb > r->max_batch) return DRL_ERR_BAD_SHAPE;
  return DRL_OK;
}

extern "C" int drl_rnd_reward(drl_rnd_t* r, const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1,
                              const float* m2, float* err_out, void* stream) {
  DRL_TRY(rnd_check(r, obs, nb, m1, m2));
  if (!err_out) return DRL_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  if (r->precision == DRL_PRECISION_BF16) {
    DRL_TRY(rnd_bf16_forward(r, obs, row_idx, nb, m1, m2, false, st));
    return rnd_bf16_mse(r, nb, err_out, nullptr, false, nullptr, st);
  }
  DRL_TRY(rnd_forward(r, obs, row_idx, nb, m1, m2, st));
  rnd_mse_kernel<<<(nb + 7) / 8 < sm_count() * 4 ? (nb + 7) / 8 : sm_count() * 4, 256, 0, st>>>(r->y, r->yh, nb, err_out, nullptr,
                                                                                            nullptr);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_rnd_learn(drl_rnd_t* r, const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1,
                             const float* m2, float* g, float* loss_out, void* stream) {
  DRL_TRY(rnd_check(r, obs, nb, m1, m2));
  if (!g || !loss_out) return DRL_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  const float* sp = r->sp;
  const int64_t* so = r->soff;
  DRL_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  DRL_CUDA(cudaMemsetAsync(g, 0, (size_t)so[12] * sizeof(float), st));
  if (r->precision == DRL_PRECISION_BF16) {
    DRL_TRY(rnd_bf16_forward(r, obs, row_idx, nb, m1, m2, true, st));
    DRL_TRY(rnd_bf16_mse(r, nb, nullptr, loss_out, true, g, st));
    return rnd_bf16_backward(r, obs, row_idx, nb, g, st);
  }
  DRL_TRY(rnd_forward(r, obs, row_idx, nb, m1, m2, st));
  rnd_mse_kernel<<<(nb + 7) / 8 < sm_count() * 4 ? (nb + 7) / 8 : sm_count() * 4, 256, 0, st>>>(r->y, r->yh, nb, nullptr, loss_out,
                                                                                            r->dyh);
  DRL_LAUNCH_CHECK();
  // student backward (random_network_distillation.py:232 loss.backward())
  DRL_TRY(simt_conv_wgrad<float>(r->dyh, r->sh[1], nullptr, nb, kStudentFc[2], g + so[10], g + so[11], st));
  DRL_TRY(simt_conv_dgrad(r->dyh, nb, kStudentFc[2], sp + so[10], r->sh[1], r->dsh[1], st, 1));
  DRL_TRY(simt_conv_wgrad<float>(r->dsh[1], r->sh[0], nullptr, nb, kStudentFc[1], g + so[8], g + so[9], st));
  DRL_TRY(simt_conv_dgrad(r->dsh[1], nb, kStudentFc[1], sp + so[8], r->sh[0], r->dsh[0], st, 1));
  DRL_TRY(simt_conv_wgrad<float>(r->dsh[0], r->sa[2], nullptr, nb, kStudentFc[0], g + so[6], g + so[7], st));
  DRL_TRY(simt_conv_dgrad(r->dsh[0], nb, kStudentFc[0], sp + so[6], r->sa[2], r->dsa[2], st, 2));
  DRL_TRY(simt_conv_wgrad<float>(r->dsa[2], r->sa[1], nullptr, nb, kTrunk[2], g + so[4], g + so[5], st));
  DRL_TRY(simt_conv_dgrad(r->dsa[2], nb, kTrunk[2], sp + so[4], r->sa[1], r->dsa[1], st, 2));
  DRL_TRY(simt_conv_wgrad<float>(r->dsa[1], r->sa[0], nullptr, nb, kTrunk[1], g + so[2], g + so[3], st));
  DRL_TRY(simt_conv_dgrad(r->dsa[1], nb, kTrunk[1], sp + so[2], r->sa[0], r->dsa[0], st, 2));
  DRL_TRY(simt_rnd_conv1_wgrad(r->dsa[0], obs, row_idx, nb, m1, m2, kTrunk[0], g + so[0], g + so[1], st));
  return DRL_OK;
}
