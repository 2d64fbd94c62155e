// Internal definition of the RND predictor handle shared by rnd.cu (entry points, fp32 CUDA-core engine) and
// umma_rnd.cu (tcgen05 bf16 engine).
#pragma once
#include "common.cuh"

namespace drl {
struct RndBf16;
}

struct drl_rnd {
  int device = 0, max_batch = 0, precision = 0;
  const float* tp = nullptr;    // teacher params (flat, reference layouts)
  const float* sp = nullptr;    // student params (flat)
  int64_t toff[9] = {0}, soff[13] = {0};
  // fp32 engine workspace
  float* ta[3] = {nullptr, nullptr, nullptr};   // teacher trunk activations
  float* sa[3] = {nullptr, nullptr, nullptr};   // student trunk activations
  float* sh[2] = {nullptr, nullptr};            // student hidden layers [nb,256]
  float *y = nullptr, *yh = nullptr, *dyh = nullptr;
  float* dsh[2] = {nullptr, nullptr};
  float* dsa[3] = {nullptr, nullptr, nullptr};
  drl::RndBf16* bf16 = nullptr;                 // tcgen05 engine (DRL_PRECISION_BF16)
};

namespace drl {

int rnd_bf16_alloc(drl_rnd* r);
void rnd_bf16_free(drl_rnd* r);
int rnd_bf16_pack(drl_rnd* r, bool teacher, bool student, cudaStream_t st);
// teacher + student forward of nb frames; fills the handle's y / yhat buffers ([nb, 512] fp32)
int rnd_bf16_forward(drl_rnd* r, const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1, const float* m2,
                     bool keep_for_backward, cudaStream_t st);
// err_out (nullable) [nb], loss_out (nullable, += mean), with_grad: also d(loss)/d(yhat) for the backward pass
int rnd_bf16_mse(drl_rnd* r, int nb, float* err_out, float* loss_out, bool with_grad, float* grads, cudaStream_t st);
int rnd_bf16_backward(drl_rnd* r, const uint8_t* obs, const int64_t* row_idx, int nb, float* grads, cudaStream_t st);

}  // namespace drl
