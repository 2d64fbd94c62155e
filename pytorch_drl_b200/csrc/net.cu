// NatureCNN agent handle: parameter layout, workspace and the orchestration of one learner step
// (forward -> fused heads/losses -> backward) and of the no-grad forward.
//
// Replaces Agent.forward (drl/agents/integration/agent.py:53-75), ActorCriticAlgo.annotate
// (drl/algos/abstract.py:371-411), PPO.compute_losses_and_metrics (drl/algos/ppo.py:172-207) and
// the backward half of PPO._optimize_losses (ppo.py:225-240).  Everything is enqueued on the
// caller's stream with no host synchronisation, so a whole step can be captured in a CUDA graph.
#include <new>
#include "net.cuh"

namespace drl {

int64_t net_param_offsets(int A, int K, int64_t* off) {
  // order of Agent.parameters(): trunk (w,b)x4, policy (w,b), value_k (w,b) ...
  const int64_t sizes[8] = {32 * 4 * 8 * 8, 32, 64 * 32 * 4 * 4, 64, 64 * 64 * 3 * 3, 64, 512 * 3136, 512};
  int64_t o = 0;
  int n = 0;
  for (int i = 0; i < 8; ++i) { if (off) off[n] = o; ++n; o += sizes[i]; }
  if (off) off[n] = o; ++n; o += (int64_t)A * 512;
  if (off) off[n] = o; ++n; o += A;
  for (int k = 0; k < K; ++k) {
    if (off) off[n] = o; ++n; o += 512;
    if (off) off[n] = o; ++n; o += 1;
  }
  if (off) off[n] = o;
  return o;
}

static void head_params(const drl_net* net, const float* params, float* grads, HeadParams& hp) {
  const int64_t* o = net->off;
  hp.wp = params + o[8]; hp.bp = params + o[9];
  hp.gwp = grads ? grads + o[8] : nullptr; hp.gbp = grads ? grads + o[9] : nullptr;
  // the bf16 engine takes the fc bias gradient from the heads kernel (the fp32 engine computes its own)
  hp.gfc_bias = (grads && net->precision == DRL_PRECISION_BF16) ? grads + o[7] : nullptr;
  for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) {
    const bool on = k < net->K;
    hp.wv[k] = on ? params + o[10 + 2 * k] : nullptr;
    hp.bv[k] = on ? params + o[11 + 2 * k] : nullptr;
    hp.gwv[k] = (on && grads) ? grads + o[10 + 2 * k] : nullptr;
    hp.gbv[k] = (on && grads) ? grads + o[11 + 2 * k] : nullptr;
  }
}

// ---- FP32 engine ----------------------------------------------------------------------------------
static int fp32_alloc(drl_net* n) {
  const size_t nb = (size_t)n->max_batch;
  const size_t e1 = nb * 20 * 20 * 32, e2 = nb * 9 * 9 * 64, e3 = nb * 7 * 7 * 64, ef = nb * 512;
  DRL_CUDA(cudaMalloc(&n->act[0], e1 * 4));
  DRL_CUDA(cudaMalloc(&n->act[1], e2 * 4));
  DRL_CUDA(cudaMalloc(&n->act[2], e3 * 4));
  DRL_CUDA(cudaMalloc(&n->dact[0], e1 * 4));
  DRL_CUDA(cudaMalloc(&n->dact[1], e2 * 4));
  DRL_CUDA(cudaMalloc(&n->dact[2], e3 * 4));
  DRL_CUDA(cudaMalloc(&n->feat, ef * 4));
  DRL_CUDA(cudaMalloc(&n->dpre, ef * 4));
  return DRL_OK;
}

static int fp32_trunk_forward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, cudaStream_t st) {
  const float* p = n->params;
  const int64_t* o = n->off;
  DRL_TRY(simt_conv_forward<uint8_t>(obs, row_idx, nb, nature_layer(0), p + o[0], p + o[1], (float*)n->act[0], st));
  DRL_TRY(simt_conv_forward<float>((const float*)n->act[0], nullptr, nb, nature_layer(1), p + o[2], p + o[3], (float*)n->act[1], st));
  DRL_TRY(simt_conv_forward<float>((const float*)n->act[1], nullptr, nb, nature_layer(2), p + o[4], p + o[5], (float*)n->act[2], st));
  DRL_TRY(simt_conv_forward<float>((const float*)n->act[2], nullptr, nb, nature_layer(3), p + o[6], p + o[7], n->feat, st));
  return DRL_OK;
}

static int fp32_trunk_backward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, float* g, cudaStream_t st) {
  const float* p = n->params;
  const int64_t* o = n->off;
  const float* dpre = (const float*)n->dpre;
  DRL_TRY(simt_conv_wgrad<float>(dpre, (const float*)n->act[2], nullptr, nb, nature_layer(3), g + o[6], g + o[7], st));
  DRL_TRY(simt_conv_dgrad(dpre, nb, nature_layer(3), p + o[6], (const float*)n->act[2], (float*)n->dact[2], st));
  DRL_TRY(simt_conv_wgrad<float>((const float*)n->dact[2], (const float*)n->act[1], nullptr, nb, nature_layer(2), g + o[4], g + o[5], st));
  DRL_TRY(simt_conv_dgrad((const float*)n->dact[2], nb, nature_layer(2), p + o[4], (const float*)n->act[1], (float*)n->dact[1], st));
  DRL_TRY(simt_conv_wgrad<float>((const float*)n->dact[1], (const float*)n->act[0], nullptr, nb, nature_layer(1), g + o[2], g + o[3], st));
  DRL_TRY(simt_conv_dgrad((const float*)n->dact[1], nb, nature_layer(1), p + o[2], (const float*)n->act[0], (float*)n->dact[0], st));
  DRL_TRY(simt_conv_wgrad<uint8_t>((const float*)n->dact[0], obs, row_idx, nb, nature_layer(0), g + o[0], g + o[1], st));
  return DRL_OK;
}

// d(features) of an external head (the DQN action-value head) -> d(pre-activation) of the fc layer in the engine's format,
// masked by the fc ReLU (feat > 0); the tcgen05 engine also needs the fc bias gradient = column sums (its heads kernel
// normally produces it).  One thread per column, 64 rows per block.
template <typename TOut>
__global__ void __launch_bounds__(512) dfeat_to_dpre_kernel(const float* __restrict__ feat, const float* __restrict__ dfeat, int nb,
                                                            TOut* __restrict__ dpre, float* __restrict__ gfc_bias) {
  const int c = threadIdx.x, r0 = blockIdx.x * 64, r1 = min(nb, r0 + 64);
  float sum = 0.f;
  for (int r = r0; r < r1; ++r) {
    const size_t e = (size_t)r * 512 + c;
    const float d = feat[e] > 0.f ? dfeat[e] : 0.f;
    sum += d;
    if constexpr (sizeof(TOut) == 2) dpre[e] = __float2bfloat16_rn(d);
    else dpre[e] = d;
  }
  if (gfc_bias) atomicAdd(gfc_bias + c, sum);
}

static int check_call(const drl_net* net, const void* obs, int nb) {
  if (!net || !obs) return DRL_ERR_BAD_ARG;
  if (!net->params) return DRL_ERR_BAD_ARG;       // drl_net_set_params first
  if (nb < 1 || nb > net->max_batch) return DRL_ERR_BAD_SHAPE;
  return DRL_OK;
}

}  // namespace drl

using namespace drl;

extern "C" int64_t drl_net_param_count(int num_actions, int num_values) {
  if (num_actions < 1 || num_actions > DRL_MAX_ACTIONS || num_values < 1 || num_values > DRL_MAX_REWARD_STREAMS) return -1;
  return net_param_offsets(num_actions, num_values, nullptr);
}

extern "C" int drl_net_param_offsets(int num_actions, int num_values, int64_t* offsets_host, int max_entries) {
  if (!offsets_host) return DRL_ERR_BAD_ARG;
  if (num_actions < 1 || num_actions > DRL_MAX_ACTIONS || num_values < 1 || num_values > DRL_MAX_REWARD_STREAMS) return DRL_ERR_BAD_SHAPE;
  if (max_entries < 2 * (5 + num_values) + 1) return DRL_ERR_BAD_SHAPE;
  net_param_offsets(num_actions, num_values, offsets_host);
  return DRL_OK;
}

extern "C" int drl_net_create(drl_net_t** out, int device, int max_batch, int num_actions, int num_values, int precision) {
  if (!out) return DRL_ERR_BAD_ARG;
  *out = nullptr;
  if (max_batch < 1 || num_actions < 1 || num_actions > DRL_MAX_ACTIONS || num_values < 1 ||
      num_values > DRL_MAX_REWARD_STREAMS) return DRL_ERR_BAD_SHAPE;
  if (precision != DRL_PRECISION_FP32 && precision != DRL_PRECISION_BF16) return DRL_ERR_BAD_ARG;
  DRL_TRY(drl_check_device(device));
  DRL_CUDA(cudaSetDevice(device));
  drl_net* n = new (std::nothrow) drl_net();
  if (!n) return DRL_ERR_CUDA;
  n->device = device; n->max_batch = max_batch; n->A = num_actions; n->K = num_values; n->precision = precision;
  n->nparams = net_param_offsets(num_actions, num_values, n->off);
  int s = precision == DRL_PRECISION_FP32 ? fp32_alloc(n) : bf16_alloc(n);
  if (s != DRL_OK) { drl_net_destroy(n); return s; }
  *out = n;
  return DRL_OK;
}

extern "C" int drl_net_destroy(drl_net_t* n) {
  if (!n) return DRL_OK;
  cudaSetDevice(n->device);
  for (int i = 0; i < 3; ++i) { cudaFree(n->act[i]); cudaFree(n->dact[i]); }
  cudaFree(n->feat); cudaFree(n->dpre);
  bf16_free(n);
  delete n;
  return DRL_OK;
}

extern "C" int drl_net_set_params(drl_net_t* net, const float* params, void* stream) {
  if (!net || !params) return DRL_ERR_BAD_ARG;
  net->params = params;
  if (net->precision == DRL_PRECISION_BF16) return bf16_pack_params(net, as_stream(stream));
  return DRL_OK;
}

extern "C" int drl_net_forward(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb,
                               float* logits, float* values, const int64_t* actions, float* logp_out,
                               float* entropy_out, void* stream) {
  DRL_TRY(check_call(net, obs, nb));
  if ((logp_out || entropy_out) && !actions && logp_out) return DRL_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  DRL_TRY(net->precision == DRL_PRECISION_FP32 ? fp32_trunk_forward(net, obs, row_idx, nb, st)
                                               : bf16_trunk_forward(net, obs, row_idx, nb, false, st));
  HeadParams hpp;
  head_params(net, net->params, nullptr, hpp);
  drl_ppo_batch_t batch{};
  batch.actions = actions;
  drl_ppo_hyper_t hp{};
  hp.num_actions = net->A; hp.num_streams = net->K;
  return heads_loss_dispatch(0, false, net->feat, hpp, row_idx, nb, batch, hp, nullptr, logits, values,
                             logp_out, entropy_out, nullptr, st);
}

// Trunk only: the 512 post-ReLU features of NatureCNN (nature_cnn.py:27-36) for heads that live outside the fused
// policy / value heads (the DQN action-value heads, csrc/dueling.cu).
extern "C" int drl_net_features(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, float* feat_out,
                                void* stream) {
  DRL_TRY(check_call(net, obs, nb));
  if (!feat_out) return DRL_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  DRL_TRY(net->precision == DRL_PRECISION_FP32 ? fp32_trunk_forward(net, obs, row_idx, nb, st)
                                               : bf16_trunk_forward(net, obs, row_idx, nb, false, st));
  DRL_CUDA(cudaMemcpyAsync(feat_out, net->feat, (size_t)nb * 512 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return DRL_OK;
}

static int ppo_common(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb,
                      const drl_ppo_batch_t* batch, const drl_ppo_hyper_t* hyper, float* grads,
                      float* losses_out, void* stream) {
  DRL_TRY(check_call(net, obs, nb));
  DRL_TRY(check_hyper(hyper));
  DRL_TRY(check_batch(batch, hyper));
  if (!losses_out) return DRL_ERR_BAD_ARG;
  if (hyper->num_actions != net->A || hyper->num_streams != net->K) return DRL_ERR_BAD_SHAPE;
  cudaStream_t st = as_stream(stream);
  const bool bf16 = net->precision == DRL_PRECISION_BF16;
  DRL_CUDA(cudaMemsetAsync(losses_out, 0, 5 * sizeof(float), st));
  if (grads) DRL_CUDA(cudaMemsetAsync(grads, 0, (size_t)net->nparams * sizeof(float), st));
  const bool heads_tc = bf16 && bf16_heads_on_tensor_cores(net);     // heads + losses as three MMAs per 128-sample tile
  DRL_TRY(bf16 ? bf16_trunk_forward(net, obs, row_idx, nb, heads_tc, st) : fp32_trunk_forward(net, obs, row_idx, nb, st));
  HeadParams hpp;
  head_params(net, net->params, grads, hpp);
  {
  ProfileScope prof("heads_loss", st);
  if (heads_tc)
    DRL_TRY(bf16_heads_loss(net, grads ? 2 : 1, hpp, row_idx, nb, *batch, *hyper, losses_out, st));
  else
    DRL_TRY(heads_loss_dispatch(grads ? 2 : 1, bf16, net->feat, hpp, row_idx, nb, *batch, *hyper, losses_out,
                                nullptr, nullptr, nullptr, nullptr, net->dpre, st));
  }
  if (grads)
    DRL_TRY(bf16 ? bf16_trunk_backward(net, obs, row_idx, nb, grads, /*overlap_comm=*/true, st)
                 : fp32_trunk_backward(net, obs, row_idx, nb, grads, st));
  if (grads && net->comm && comm_world(net->comm) > 1) {
    // bf16 engine: the fc + heads range (queued after the fc weight gradient) and the conv2 + conv3 range (after conv2's weight
    // gradient) are already in flight; here conv1's.  fp32 engine: the whole buffer.
    const int64_t lo = 0, hi = bf16 ? net->off[2] : net->nparams;
    DRL_TRY(comm_allreduce_after(net->comm, grads, lo, hi, st));
    DRL_TRY(comm_join(net->comm, st));
  }
  return DRL_OK;
}

// Vector-Jacobian product of the whole network: gradients of sum(dlogits * logits) + sum(dvalues * values) with
// respect to every parameter, for callers that build their own loss on top of drl_net_forward (a differentiable
// Agent.forward, agent.py:53-75).  The activations are recomputed (the handle keeps no tape between calls).
extern "C" int drl_net_backward(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, const float* dlogits,
                                const float* dvalues, float* grads, void* stream) {
  DRL_TRY(check_call(net, obs, nb));
  if (!dlogits || !dvalues || !grads) return DRL_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  const bool bf16 = net->precision == DRL_PRECISION_BF16;
  DRL_CUDA(cudaMemsetAsync(grads, 0, (size_t)net->nparams * sizeof(float), st));
  DRL_TRY(bf16 ? bf16_trunk_forward(net, obs, row_idx, nb, false, st) : fp32_trunk_forward(net, obs, row_idx, nb, st));
  HeadParams hpp;
  head_params(net, net->params, grads, hpp);
  drl_ppo_batch_t batch{};
  drl_ppo_hyper_t hp{};
  hp.num_actions = net->A; hp.num_streams = net->K;
  DRL_TRY(heads_loss_dispatch(3, bf16, net->feat, hpp, row_idx, nb, batch, hp, nullptr, const_cast<float*>(dlogits),
                              const_cast<float*>(dvalues), nullptr, nullptr, net->dpre, st));
  // no collective here: the caller owns `grads` (a scratch buffer of the autograd path) and decides when to exchange it
  DRL_TRY(bf16 ? bf16_trunk_backward(net, obs, row_idx, nb, grads, /*overlap_comm=*/false, st)
               : fp32_trunk_backward(net, obs, row_idx, nb, grads, st));
  return DRL_OK;
}

extern "C" int drl_net_backward_features(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, const float* dfeat,
                                         float* grads, int reuse_forward, void* stream) {
  DRL_TRY(check_call(net, obs, nb));
  if (!dfeat || !grads) return DRL_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  const bool bf16 = net->precision == DRL_PRECISION_BF16;
  DRL_CUDA(cudaMemsetAsync(grads, 0, (size_t)net->nparams * sizeof(float), st));
  // reuse_forward: the caller's LAST call on this handle was drl_net_features on exactly (obs, row_idx, nb): the activations are live
  if (!reuse_forward) DRL_TRY(bf16 ? bf16_trunk_forward(net, obs, row_idx, nb, false, st) : fp32_trunk_forward(net, obs, row_idx, nb, st));
  if (bf16)
    dfeat_to_dpre_kernel<__nv_bfloat16><<<(nb + 63) / 64, 512, 0, st>>>(net->feat, dfeat, nb, reinterpret_cast<__nv_bfloat16*>(net->dpre),
                                                                       grads + net->off[7]);
  else
    dfeat_to_dpre_kernel<float><<<(nb + 63) / 64, 512, 0, st>>>(net->feat, dfeat, nb, reinterpret_cast<float*>(net->dpre), nullptr);
  DRL_LAUNCH_CHECK();
  DRL_TRY(bf16 ? bf16_trunk_backward(net, obs, row_idx, nb, grads, /*overlap_comm=*/false, st)
               : fp32_trunk_backward(net, obs, row_idx, nb, grads, st));
  return DRL_OK;
}

extern "C" int drl_net_set_comm(drl_net_t* net, drl_comm_t* comm) {
  if (!net) return DRL_ERR_BAD_ARG;
  net->comm = comm;
  return DRL_OK;
}

extern "C" int drl_net_ppo_step(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb,
                                const drl_ppo_batch_t* batch, const drl_ppo_hyper_t* hyper, float* grads,
                                float* losses_out, void* stream) {
  if (!grads) return DRL_ERR_BAD_ARG;
  return ppo_common(net, obs, row_idx, nb, batch, hyper, grads, losses_out, stream);
}

extern "C" int drl_net_ppo_metrics(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb,
                                   const drl_ppo_batch_t* batch, const drl_ppo_hyper_t* hyper,
                                   float* losses_out, void* stream) {
  return ppo_common(net, obs, row_idx, nb, batch, hyper, nullptr, losses_out, stream);
}

extern "C" int drl_net_activation(drl_net_t* net, int which, const void** ptr, int64_t* elems, int* dtype_bytes) {
  if (!net || !ptr || which < 1 || which > 4) return DRL_ERR_BAD_ARG;
  const int64_t per[4] = {20 * 20 * 32, 9 * 9 * 64, 7 * 7 * 64, 512};
  if (which == 4) { *ptr = net->feat; if (dtype_bytes) *dtype_bytes = 4; }
  else { *ptr = net->act[which - 1]; if (dtype_bytes) *dtype_bytes = net->precision == DRL_PRECISION_FP32 ? 4 : 2; }
  if (elems) *elems = per[which - 1] * net->max_batch;
  return DRL_OK;
}

extern "C" int drl_net_adam_step(drl_net_t* net, float* params, const float* grads, float* exp_avg,
                                 float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                                 int64_t step, float grad_scale, void* stream) {
  if (!net) return DRL_ERR_BAD_ARG;
  {
  ProfileScope prof("adam", as_stream(stream));
  DRL_TRY(adam_launch(params, grads, exp_avg, exp_avg_sq, net->nparams, lr, beta1, beta2, eps, step,
                      grad_scale, as_stream(stream)));
  }
  net->params = params;
  if (net->precision == DRL_PRECISION_BF16) return bf16_pack_params(net, as_stream(stream));
  return DRL_OK;
}
