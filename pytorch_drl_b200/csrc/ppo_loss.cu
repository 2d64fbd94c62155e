// Stand-alone fused PPO loss kernel on given logits / values (C-ABI drl_ppo_loss_fwd_bwd).
// One thread per sample, the minibatch gather (drl/utils/nested.py:6-12) folded in through
// row_idx, five scalar reductions by warp shuffles and one set of atomics per block.
// HBM-bound: algorithmic bytes 8A + 12 + 20K per sample.
#include "ppo_math.cuh"

namespace drl {

// AX / KX: exact action / stream counts for the common Atari heads (0 = dynamic), see ppo_math.cuh.  The kernel is
// instruction-issue bound (ncu: ~590 warp instructions per sample at A = 6 before the exact-count instantiations), not
// access-pattern bound: staging the logits tile through shared memory for coalesced copies made it slower (119 -> 199 us
// at 4 M samples, profiles/r01_notes.md).
template <int MAXA, int AX = 0, int KX = 0>
__global__ void __launch_bounds__(256) ppo_loss_kernel(
    const float* __restrict__ logits, const float* __restrict__ values,
    const int64_t* __restrict__ row_idx, int nb, drl_ppo_batch_t batch, drl_ppo_hyper_t hp,
    float inv_nb, float* __restrict__ losses_out, float* __restrict__ logp_out,
    float* __restrict__ ent_out, float* __restrict__ dlogits, float* __restrict__ dvalues) {
  __shared__ float red[128];
  const int A = AX ? AX : hp.num_actions, K = KX ? KX : hp.num_streams;
  PpoSums acc{0.f, 0.f, 0.f, 0.f};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nb; i += gridDim.x * blockDim.x) {
    float logit[MAXA], dlogit[MAXA], v_new[DRL_MAX_REWARD_STREAMS], dv[DRL_MAX_REWARD_STREAMS];
#pragma unroll
    for (int j = 0; j < MAXA; ++j) if (j < A) logit[j] = logits[(size_t)i * A + j];
#pragma unroll
    for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) if (k < K) v_new[k] = values[(size_t)i * K + k];
    PpoSampleIn in;
    load_sample(batch, K, row_idx ? row_idx[i] : (int64_t)i, in);
    PpoSums s;
    float logp, ent;
    ppo_sample<MAXA, AX, KX>(logit, v_new, in, hp, inv_nb, dlogits != nullptr, dlogit, dv, s, logp, ent);
    acc.surr += s.surr; acc.ent += s.ent; acc.clip += s.clip; acc.vf += s.vf;
    if (logp_out) logp_out[i] = logp;
    if (ent_out) ent_out[i] = ent;
    if (dlogits) {
#pragma unroll
      for (int j = 0; j < MAXA; ++j) if (j < A) dlogits[(size_t)i * A + j] = dlogit[j];
    }
    if (dvalues) {
#pragma unroll
      for (int k = 0; k < DRL_MAX_REWARD_STREAMS; ++k) if (k < K) dvalues[(size_t)i * K + k] = dv[k];
    }
  }
  PpoSums tot = block_reduce_sums(acc, red);
  if (threadIdx.x == 0) ppo_accumulate_losses(tot, hp, inv_nb, losses_out);
}

int check_hyper(const drl_ppo_hyper_t* hp) {
  if (!hp) return DRL_ERR_BAD_ARG;
  if (hp->num_actions < 1 || hp->num_actions > DRL_MAX_ACTIONS) return DRL_ERR_BAD_SHAPE;
  if (hp->num_streams < 1 || hp->num_streams > DRL_MAX_REWARD_STREAMS) return DRL_ERR_BAD_SHAPE;
  if (hp->vf_loss_kind != 0 && hp->vf_loss_kind != 1) return DRL_ERR_BAD_ARG;
  return DRL_OK;
}

int check_batch(const drl_ppo_batch_t* b, const drl_ppo_hyper_t* hp) {
  if (!b || !b->actions || !b->logp_old) return DRL_ERR_BAD_ARG;
  for (int k = 0; k < hp->num_streams; ++k) {
    if (!b->adv[k] || !b->vtarg[k]) return DRL_ERR_BAD_ARG;
    if (hp->vf_loss_clipping && !b->v_old[k]) return DRL_ERR_BAD_ARG;
  }
  return DRL_OK;
}

}  // namespace drl

using namespace drl;

extern "C" int drl_ppo_loss_fwd_bwd(const float* logits, const float* values,
                                    const int64_t* row_idx, int nb, const drl_ppo_batch_t* batch,
                                    const drl_ppo_hyper_t* hyper, float* losses_out,
                                    float* logp_out, float* entropy_out, float* dlogits,
                                    float* dvalues, void* stream) {
  DRL_TRY(check_hyper(hyper));
  DRL_TRY(check_batch(batch, hyper));
  if (!logits || !values || !losses_out) return DRL_ERR_BAD_ARG;
  if ((dlogits == nullptr) != (dvalues == nullptr)) return DRL_ERR_BAD_ARG;
  if (nb < 1) return DRL_ERR_BAD_SHAPE;
  cudaStream_t st = as_stream(stream);
  DRL_CUDA(cudaMemsetAsync(losses_out, 0, 5 * sizeof(float), st));
  const int threads = 256;
  int blocks = (nb + threads - 1) / threads;
  const int maxb = sm_count() * 8;
  if (blocks > maxb) blocks = maxb;
  const float inv_nb = 1.0f / (float)nb;
  const int A = hyper->num_actions;
#define LAUNCH(MAXA, AX, KX)                                                                                \
  ppo_loss_kernel<MAXA, AX, KX><<<blocks, threads, 0, st>>>(logits, values, row_idx, nb, *batch, *hyper, inv_nb, \
                                                            losses_out, logp_out, entropy_out, dlogits, dvalues)
  const int K = hyper->num_streams;
  if (A == 6 && K == 1) LAUNCH(8, 6, 1);
  else if (A == 4 && K == 1) LAUNCH(8, 4, 1);
  else if (A == 9 && K == 1) LAUNCH(18, 9, 1);
  else if (A == 18 && K == 1) LAUNCH(18, 18, 1);
  else if (A == 18 && K == 2) LAUNCH(18, 18, 2);
  else if (A <= 8) LAUNCH(8, 0, 0);
  else if (A <= 18) LAUNCH(18, 0, 0);
  else LAUNCH(32, 0, 0);
#undef LAUNCH
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
