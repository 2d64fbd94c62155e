// Internal definition of the NatureCNN agent handle shared by net.cu and the precision engines.
#pragma once
#include "common.cuh"
#include "simt_conv.cuh"

namespace drl {

struct HeadParams;
struct Bf16Engine;

}  // namespace drl

struct drl_comm;
struct drl_net {
  int device = 0, max_batch = 0, A = 0, K = 0, precision = 0;
  drl_comm* comm = nullptr;            // if attached: ppo_step all-reduces the gradients, overlapped with the backward
  int64_t nparams = 0;
  int64_t off[8 + 2 + 2 * DRL_MAX_REWARD_STREAMS + 1] = {0};
  const float* params = nullptr;       // caller-owned flat fp32 parameters (reference layouts)
  void* act[3] = {nullptr, nullptr, nullptr};    // conv activations, NHWC (fp32 or bf16)
  void* dact[3] = {nullptr, nullptr, nullptr};   // their gradients (pre-activation, ReLU-masked)
  float* feat = nullptr;               // [nb, 512] fp32 features (post-ReLU)
  void* dpre = nullptr;                // [nb, 512] d composite / d (fc pre-activation)
  drl::Bf16Engine* bf16 = nullptr;
};

namespace drl {

struct HeadParams {            // pointers into the flat parameter / gradient buffers
  const float* wp; const float* bp;
  const float* wv[DRL_MAX_REWARD_STREAMS]; const float* bv[DRL_MAX_REWARD_STREAMS];
  float* gwp; float* gbp;
  float* gwv[DRL_MAX_REWARD_STREAMS]; float* gbv[DRL_MAX_REWARD_STREAMS];
  float* gfc_bias;               // if set: += column sums of d(feat) (the fc layer's bias gradient), fused
};

int64_t net_param_offsets(int A, int K, int64_t* off);
int check_hyper(const drl_ppo_hyper_t* hp);
int check_batch(const drl_ppo_batch_t* b, const drl_ppo_hyper_t* hp);
int heads_loss_dispatch(int mode, bool dpre_is_bf16, const float* feat, const HeadParams& hpp,
                        const int64_t* row_idx, int nb, const drl_ppo_batch_t& batch,
                        const drl_ppo_hyper_t& hp, float* losses_out, float* logits_out,
                        float* values_out, float* logp_out, float* ent_out, void* dpre,
                        cudaStream_t st);
int adam_launch(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                double lr, double beta1, double beta2, double eps, int64_t step, float grad_scale,
                cudaStream_t st);

// tcgen05 bf16 engine (umma_net.cu)
int bf16_alloc(drl_net* n);
void bf16_free(drl_net* n);
int bf16_pack_params(drl_net* n, cudaStream_t st);
// feat_bf16: the fc layer writes bf16 features for the tensor-core heads kernel (learner step / metrics pass) instead of fp32 `feat`
int bf16_trunk_forward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, bool feat_bf16, cudaStream_t st);
bool bf16_heads_on_tensor_cores(const drl_net* n);
int bf16_heads_loss(drl_net* n, int mode, const HeadParams& hpp, const int64_t* row_idx, int nb, const drl_ppo_batch_t& batch,
                    const drl_ppo_hyper_t& hp, float* losses_out, cudaStream_t st);
int comm_allreduce_after(drl_comm* c, float* flat, int64_t lo, int64_t hi, cudaStream_t producer);
int comm_join(drl_comm* c, cudaStream_t consumer);
int comm_world(const drl_comm* c);
// overlap_comm: start the all-reduce of the fc + heads range on the attached communicator's side stream as soon as it is final
// (only the learner step does: it joins the exchange afterwards)
int bf16_trunk_backward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, float* grads, bool overlap_comm, cudaStream_t st);

}  // namespace drl
