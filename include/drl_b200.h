/*
 * drl_b200.h — C-ABI of the B200-native learner hot path of lucaslingle/pytorch_drl.
 *
 * The reference has no FFI of its own: the path sits behind Python classes selected by name
 * from YAML.  This header is the boundary a replacement exports underneath
 * those classes: plain pointers and sizes, caller-owned device buffers, an explicit CUDA stream
 * (passed as void*), int status codes, never a throw/abort across the boundary.  The Python
 * shims in `pytorch_drl_b200/` bind it with ctypes and map status codes to ValueError /
 * RuntimeError the way the reference raises them.  INTEGRATION.md shows the binding.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference
 * repository root).  All pointers are DEVICE pointers unless the name ends in `_host`.
 * All calls are asynchronous with respect to the host and ordered on `stream`.
 * A handle is not thread-safe; distinct handles are independent.
 */
#ifndef DRL_B200_H_
#define DRL_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif
#ifdef __cplusplus
extern "C" {
#endif

#define DRL_B200_ABI_VERSION 1

/* ---- status codes (Python shim: 1,2,3 -> ValueError; others -> RuntimeError) ---------------- */
enum {
  DRL_OK = 0,
  DRL_ERR_BAD_SHAPE = 1,    /* e.g. T % B != 0 (drl/algos/common/samplers.py:51), N<=0, A out of range */
  DRL_ERR_BAD_ARG = 2,      /* null pointer, unknown enum value (drl/algos/common/credit_assignment.py:25) */
  DRL_ERR_MISALIGNED = 3,   /* a pointer violates the stated alignment */
  DRL_ERR_CUDA = 4,         /* a CUDA runtime/driver call failed; see drl_last_cuda_error() */
  DRL_ERR_NO_DEVICE = 5,    /* no sm_100 device: there is NO CPU fallback */
  DRL_ERR_UNSUPPORTED = 6,  /* valid request this build does not implement */
  DRL_ERR_NCCL = 7
};

/* DRL_API: every symbol declared in this header is exported (default visibility). */
int drl_abi_version(void);
const char* drl_status_string(int status);
/* cudaGetErrorString of the last CUDA failure seen by this library on this thread ("" if none). */
const char* drl_last_cuda_error(void);
/* 0 iff device `device` exists and is compute capability 10.x. */
int drl_check_device(int device);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
long long drl_launch_count(void);
/* Per-launch-site CUDA-event timing on the launching stream.  drl_profile_read synchronises the
 * device, returns for each site seen since the last read its name (48-char slots), summed
 * milliseconds and launch count, and clears the records. */
int drl_profile_enable(int on);
/* Kernel-variant switches for A/B measurements (keys 6-8: CTA-pair / epilogue-warpgroup variants of the fc and conv
 * kernels, heads-kernel occupancy; see csrc/umma_net.cu).  Defaults are the measured-fastest variants. */
int drl_debug_set(int key, int value);
int drl_profile_read(int max_sites, char* names_out, float* total_ms, int* counts, int* n_sites);

/* =============================================================================================
 * Credit assignment — replaces GAE.call (drl/algos/common/credit_assignment.py:195-214), the
 * vtarg/standardize tail of ActorCriticAlgo.credit_assignment (drl/algos/abstract.py:427-439)
 * and standardize (drl/utils/stats.py:4-17), for n_envs independent trajectories at once.
 *
 *   rewards [n_envs, ld_r]  first T entries used      vpreds [n_envs, ld_v]  first T+1 used
 *   dones   [n_envs, ld_r]  (float 0/1)               adv/vtarg [n_envs, ld_o] first T written
 *   nt = use_dones ? 1-d_t : 1;  delta_t = -V_t + r_t + nt*gamma*V_{t+1};
 *   A_t = delta_t + nt*gamma*lambda*A_{t+1};  vtarg = A + V[:T];  if standardize:
 *   A = (A-mean(A))/std_unbiased(A) per env (the reference's ranks are envs).
 * One warp per env, reverse affine scan, float4 loads when T%4==0 and rows are 16B aligned.
 * ============================================================================================= */
int drl_gae_f32(const float* rewards, const float* vpreds, const float* dones,
                int n_envs, int T, int ld_r, int ld_v, int ld_o,
                float gamma, float lambda, int use_dones, int standardize,
                float* adv_out, float* vtarg_out, void* stream);

/* standardize alone (drl/utils/stats.py:4-17) on n_rows rows of length T. */
int drl_standardize_f32(const float* x, int n_rows, int T, int ld, float* y, void* stream);

/* =============================================================================================
 * Fused PPO losses — replaces Categorical.log_prob/entropy (drl/algos/abstract.py:401-402),
 * ppo_policy_entropy_bonus / ppo_policy_surrogate_objective / ppo_vf_loss
 * (drl/algos/ppo.py:310-417), the composite (ppo.py:195-200) and their autograd backward
 * w.r.t. logits and values, in one reduction kernel (warp shuffles + one atomic per block).
 * ============================================================================================= */
#define DRL_MAX_REWARD_STREAMS 4
#define DRL_MAX_ACTIONS 32

typedef struct {
  int num_actions;                              /* A */
  int num_streams;                              /* K reward streams, ordered like reward_weights */
  float reward_weights[DRL_MAX_REWARD_STREAMS]; /* ppo.py:358-359 */
  float clip_param;                             /* LinearSchedule value, schedules.py:18-29 */
  float ent_coef;
  float vf_coef;                                /* ppo.py:198 */
  int vf_loss_clipping;                         /* ppo.py:403-409 */
  int vf_simple_weighting;                      /* ppo.py:413-416 */
  int vf_loss_kind;                             /* 0 MSELoss, 1 SmoothL1Loss (losses.py:6-18) */
} drl_ppo_hyper_t;

/* Per-sample tensors indexed through row_idx (NULL = identity): sample i of the minibatch reads
 * row r = row_idx[i] of actions/logp_old/adv/v_old/vtarg — the gather of slice_nested_tensor
 * (drl/utils/nested.py:6-12) folded into the kernel.  logits [nb, A] and values [nb, K] are the
 * NEW predictions (dense, not indexed).  adv/v_old/vtarg: K device pointers passed by value.
 * losses_out[5] (device) = policy_loss, value_loss, composite_loss, meanent, clipfrac
 * (ppo.py:201-207); it is zeroed by the call.  dlogits [nb, A] / dvalues [nb, K] may be NULL
 * (metrics pass, ppo.py:265-266); they hold d composite / d(.) with the 1/nb of the means. */
typedef struct {
  const int64_t* actions;
  const float* logp_old;
  const float* adv[DRL_MAX_REWARD_STREAMS];
  const float* v_old[DRL_MAX_REWARD_STREAMS];
  const float* vtarg[DRL_MAX_REWARD_STREAMS];
} drl_ppo_batch_t;

int drl_ppo_loss_fwd_bwd(const float* logits, const float* values, const int64_t* row_idx, int nb,
                         const drl_ppo_batch_t* batch, const drl_ppo_hyper_t* hyper,
                         float* losses_out, float* logp_out, float* entropy_out,
                         float* dlogits, float* dvalues, void* stream);

/* =============================================================================================
 * NatureCNN agent — replaces Agent.forward (drl/agents/integration/agent.py:53-75) with
 * ToChannelMajor (preprocessing/vision.py:4-9), NatureCNN (architectures/stateless/
 * nature_cnn.py:27-36), CategoricalPolicyHead logits (heads/policy_heads.py:107) and
 * SimpleValueHead (heads/value_heads.py:64), plus the autograd backward and the minibatch step
 * of PPO._optimize_losses (drl/algos/ppo.py:225-240).
 *
 * Parameters live in ONE caller-owned flat fp32 buffer in the reference's own order and layouts
 * (Agent.parameters(): conv OIHW weights, fc [512, c*49+h*7+w], heads [A,512],[1,512]; see
 * drl_net_param_count / drl_net_param_offsets), so torch views of it are state_dict compatible
 * (drl/utils/checkpointing.py:64).  Gradients are written to a second buffer of the same layout.
 * Observations are uint8 NHWC [rows, 84, 84, 4] as the env emits them before ScaleObservations
 * (scale_observations.py:35): the x*(1/255) is fused into conv1.
 * ============================================================================================= */
typedef struct drl_net drl_net_t;

enum { DRL_PRECISION_FP32 = 0,   /* CUDA-core fp32 implicit GEMM: the strict-parity mode          */
       DRL_PRECISION_BF16 = 1 }; /* tcgen05 bf16 implicit GEMM, fp32 accumulate: throughput mode */

int drl_net_create(drl_net_t** out, int device, int max_batch, int num_actions, int num_values,
                   int precision);
int drl_net_destroy(drl_net_t* net);
/* number of fp32 parameters (1 687 719 for A=6,K=1) and the 2*(5+K) tensor offsets in order. */
int64_t drl_net_param_count(int num_actions, int num_values);
int drl_net_param_offsets(int num_actions, int num_values, int64_t* offsets_host, int max_entries);

/* (Re)bind the flat parameter buffer; in bf16 mode also refreshes the packed bf16 operand copies.
 * Must be called after the caller changed parameters by other means than drl_net_adam_step. */
int drl_net_set_params(drl_net_t* net, const float* params, void* stream);

/* No-grad forward (annotate, drl/algos/abstract.py:383-411; actor, rollout.py:284-287):
 * logits [nb, A], values [nb, K]; logp/entropy of `actions` (indexed by row_idx) if non-NULL. */
int drl_net_forward(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb,
                    float* logits, float* values, const int64_t* actions, float* logp_out,
                    float* entropy_out, void* stream);

/* Trunk only: feat_out [nb, 512] = the post-ReLU NatureCNN features (nature_cnn.py:27-36) of the minibatch, for heads
 * outside the fused policy / value heads (the DQN action-value heads: drl_dueling_forward_f32). */
int drl_net_features(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, float* feat_out, void* stream);

/* One learner step without the optimiser: forward, fused losses, backward.  grads (flat, same
 * layout as params) is OVERWRITTEN with d composite / d params, where composite is the mean over
 * the nb samples (== mean over ranks of per-rank means for equal shards).
 * losses_out[5] as in drl_ppo_loss_fwd_bwd. */
int drl_net_ppo_step(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb,
                     const drl_ppo_batch_t* batch, const drl_ppo_hyper_t* hyper,
                     float* grads, float* losses_out, void* stream);

/* Full-rollout metrics pass (ppo.py:265-266): forward + losses, no backward. */
int drl_net_ppo_metrics(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb,
                        const drl_ppo_batch_t* batch, const drl_ppo_hyper_t* hyper,
                        float* losses_out, void* stream);

/* debug / parity access to the last step's intermediate activations (device pointers owned by
 * the handle; layout NHWC; dtype fp32 in FP32 mode, bf16 in BF16 mode). which: 1..3 conv
 * activations, 4 features [nb,512] (always fp32). */
int drl_net_activation(drl_net_t* net, int which, const void** ptr, int64_t* elems, int* dtype_bytes);
/* device-to-device copy on `stream` (lets a binding read handle-owned buffers). */
int drl_memcpy_d2d(void* dst, const void* src, size_t bytes, void* stream);
/* Copy done by a kernel (no copy engine): src may be pinned host memory; bytes and both pointers multiples of 4.  Used for the
 * per-epoch minibatch indices so that they never queue behind a bulk rollout upload. */
int drl_copy_by_kernel(void* dst, const void* src, size_t bytes, void* stream);

/* =============================================================================================
 * Optimiser — replaces torch.optim.Adam.step as configured by drl/utils/optimization.py:9-35
 * (models_dir/atari_ppo/config.yaml:76-78: betas, eps=1e-5; no amsgrad, no weight decay) on a
 * flat buffer.  grad_scale multiplies the gradient first (1/world after a SUM allreduce).
 * `step` is the 1-based update count.  One vectorised pass: 28 B/param of HBM traffic.
 * Hyper-parameters are doubles because torch forms 1-beta and lr/bias_correction in double.
 * ============================================================================================= */
int drl_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  double lr, double beta1, double beta2, double eps, int64_t step, float grad_scale,
                  void* stream);
/* Adam on the handle's parameters, also refreshing the packed operand copies (BF16 mode). */
int drl_net_adam_step(drl_net_t* net, float* params, const float* grads, float* exp_avg,
                      float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                      int64_t step, float grad_scale, void* stream);

/* =============================================================================================
 * n-step credit assignment — replaces NStepAdvantageEstimator.call and
 * SimpleDiscreteBellmanOptimalityOperator.call (drl/algos/common/credit_assignment.py:225-249, 276-305).
 * n = extra_steps + 1; rewards/dones [n_envs, >= T+n-1] with row stride ld_r, vpreds [n_envs, >= T+n] (ld_v),
 * q tables [n_envs, >= T+n, A] with row stride ld_q (in time steps).  Bit-identical to the reference.
 * ============================================================================================= */
/* adv[e,t] = r_t + nt_t g r_{t+1} + ... + (prod nt) g^n V(s_{t+n}) - V(s_t) */
int drl_nstep_advantages_f32(const float* rewards, const float* vpreds, const float* dones, int n_envs, int T,
                             int extra_steps, int64_t ld_r, int64_t ld_v, int64_t ld_out, float gamma, int use_dones,
                             float* adv_out, void* stream);
/* out[e,t] = n-step return bootstrapped with Qtgt(s_{t+n}, argmax_a Q*(s_{t+n}, a)), Q* = qpreds if double_q else
 * tgt_qpreds (first maximum on ties, like torch.argmax). */
int drl_bellman_backups_f32(const float* rewards, const float* qpreds, const float* tgt_qpreds, const float* dones,
                            int n_envs, int T, int extra_steps, int num_actions, int64_t ld_r, int64_t ld_q,
                            int64_t ld_out, float gamma, int use_dones, int double_q, float* out, void* stream);

/* =============================================================================================
 * DQN head blocks — replace DuelingArchitecture.forward behind SimpleDiscreteActionValueHead.forward
 * (drl/agents/architectures/stateless/dueling.py:50-57, drl/agents/heads/action_value_heads.py:124-140) and the
 * arithmetic of EpsilonGreedyCategoricalPolicyHead.forward (drl/agents/heads/policy_heads.py:219-224).
 * ============================================================================================= */
/* q[nb, A] = (adv - mean(adv)) + val from features [nb, in_dim]; weights in torch nn.Linear layout ([out, in]):
 * w0 [50, in_dim], wa1 / wv1 [25, 50], wa2 [A, 25], wv2 [1, 25]; both streams end in a ReLU like the reference's MLP
 * (mlp.py:32-41).  widening must be 1, in_dim <= 512, A <= DRL_MAX_ACTIONS.  Forward only. */
int drl_dueling_forward_f32(const float* feat, int nb, int in_dim, int num_actions, int widening, const float* w0,
                            const float* b0, const float* wa1, const float* ba1, const float* wa2, const float* ba2,
                            const float* wv1, const float* bv1, const float* wv2, const float* bv2, float* q_out,
                            void* stream);
/* DQN learner step on that head (BASELINE config "Double/Dueling DQN with prioritized replay").  The reference ships the head
 * (dueling.py:9-57, action_value_heads.py:85-140) and the double-Q Bellman operator (credit_assignment.py:252-305) but no DQN
 * algorithm; this entry is the loss a learner puts on them and its gradient, for a batch of sampled transitions:
 *   y_i = rewards_i + (1 - dones_i) * gamma * q_next_target[i, argmax_a q_sel[i, a]]     q_sel = q_next_online if double_q else q_next_target
 *   td_i = Q(s_i, actions_i) - y_i ;  loss = mean_i is_weights_i * crit(td_i) ;  crit = SmoothL1 (beta 1) if smooth_l1 else squared error
 * head_params / head_grads: ONE flat fp32 buffer each in the reference module's parameters() order
 *   [w0 50 x in_dim | b0 50 | wa1 25 x 50 | ba1 25 | wa2 A x 25 | ba2 A | wv1 25 x 50 | bv1 25 | wv2 25 | bv2 1];
 * head_grads and loss_out[1] are ACCUMULATED into (caller zeroes them), dfeat_out [nb, in_dim] = d loss / d features is
 * overwritten (feed it to drl_net_backward_features), q_out [nb, A] / y_out [nb] are optional, td_out [nb] feeds the priority
 * update (drl_sumtree_update), workspace: nb * 50 floats.  is_weights may be NULL (all ones). */
int drl_dueling_loss_backward_f32(const float* feat, int nb, int in_dim, int num_actions, int widening,
                                  const float* head_params, const int64_t* actions, const float* rewards, const float* dones,
                                  const float* q_next_online, const float* q_next_target, const float* is_weights, float gamma,
                                  int use_dones, int double_q, int smooth_l1, float* q_out, float* td_out, float* y_out,
                                  float* loss_out, float* dfeat_out, float* head_grads, float* workspace, void* stream);
/* probs[nb, A] = (1 - eps) * one_hot(argmax_a q) + eps / A (first maximum on ties, like torch.argmax), bit-identical
 * to the reference's float32 expression; greedy_out (nullable) receives the argmax.  eps outside [0, 1] is refused. */
int drl_epsilon_greedy_probs_f32(const float* q, int nb, int num_actions, double epsilon, float* probs_out,
                                 int64_t* greedy_out, void* stream);

/* =============================================================================================
 * Reward normaliser — replaces ReturnAcc and the per-step part of NormalizeRewardWrapper.step
 * (drl/envs/wrappers/stateful/normalize_reward.py:20-100, 171-194) for n_envs environments at once.
 * ============================================================================================= */
/* Feed T new (reward, done) steps of every environment (rewards[e*r_ld + t], dones likewise, 0/1) into its
 * un-synchronised accumulator.  State (caller-allocated, zero-initialised): win_r / win_d float[2*trace_len][n_envs]
 * (time-major window of pending pairs), count int[n_envs], steps / m1 / m2 float[n_envs].
 * trace_len = int(5 / (1 - gamma)) (normalize_reward.py:27). */
int drl_return_acc_update_f32(const float* rewards, int64_t r_ld, const float* dones, int64_t d_ld, int n_envs, int T,
                              float gamma, int use_dones, int trace_len, float* win_r, float* win_d, int* count,
                              float* steps, float* m1, float* m2, void* stream);
/* y[i] = 0 if steps == 0 else clip(x[i] * rsqrt(m2 - m1^2 + eps), lo, hi), (steps, m1, m2) = synced_moments[0..2]
 * on the device (ReturnAcc.forward, shift = False: normalize_reward.py:97-100, normalize.py:150-162). */
int drl_reward_norm_apply_f32(const float* x, int64_t n, const float* synced_moments, float eps, float clip_lo,
                              float clip_hi, float* y, void* stream);

/* =============================================================================================
 * Observation normaliser — replaces Normalizer.forward / update
 * (drl/envs/wrappers/stateful/normalize.py:129-162, 105-127).
 * ============================================================================================= */
/* y[i, j] = clip((x[i, j] - m1[j]) * rsqrt(m2[j] - m1[j]^2 + eps), lo, hi); x is n x d */
int drl_normalizer_apply_f32(const float* x, const float* m1, const float* m2, int64_t n, int64_t d,
                             float eps, float clip_lo, float clip_hi, float* y, void* stream);
/* Running-moment update with `count` items x[count, d] applied in order; steps_before is the
 * normaliser's step count before the first item. */
int drl_normalizer_update_f32(float* m1, float* m2, const float* x, int64_t count, int64_t d,
                              double steps_before, void* stream);

/* =============================================================================================
 * RND predictor — replaces the arithmetic of RandomNetworkDistillationWrapper.learn and .step
 * (drl/envs/wrappers/stateful/intrinsic/random_network_distillation.py:209-235, 186-201) with
 * _TeacherNetwork / _StudentNetwork (same file :22-87).  Flat parameter buffers in
 * `named_parameters()` order, reference layouts.
 * ============================================================================================= */
typedef struct drl_rnd drl_rnd_t;
int drl_rnd_create(drl_rnd_t** out, int device, int max_batch, int widening, int precision);
int drl_rnd_destroy(drl_rnd_t* rnd);
int64_t drl_rnd_param_count(int student, int widening);
/* precision: DRL_PRECISION_FP32 (CUDA cores, strict parity) or DRL_PRECISION_BF16 (tcgen05: sample pairs packed along the
 * channels of block-diagonal implicit GEMMs, csrc/umma_rnd.cu).  (Re)binds the flat parameter buffers and, in BF16 mode,
 * refreshes the packed operand copies of both networks. */
int drl_rnd_set_params(drl_rnd_t* rnd, const float* teacher_params, const float* student_params,
                       void* stream);
/* Adam on the student's flat parameters (rnd_optimizer.step(), random_network_distillation.py:233), also refreshing the
 * student's packed operand copies (BF16 mode).  Arguments as drl_adam_step. */
int drl_rnd_adam_step(drl_rnd_t* rnd, float* student_params, const float* grads, float* exp_avg, float* exp_avg_sq,
                      double lr, double beta1, double beta2, double eps, int64_t step, float grad_scale, void* stream);
/* Intrinsic reward (step, :193-196): err[i] = mean_d (teacher - student)^2 of the normalised,
 * clipped LAST channel of uint8 frame row_idx[i]; no gradient. */
int drl_rnd_reward(drl_rnd_t* rnd, const uint8_t* obs, const int64_t* row_idx, int nb,
                   const float* m1, const float* m2, float* err_out, void* stream);
/* Predictor loss and student gradients (learn, :225-232): loss_out[0] = mean_b mean_d (y-yh)^2;
 * student_grads (flat) overwritten. */
int drl_rnd_learn(drl_rnd_t* rnd, const uint8_t* obs, const int64_t* row_idx, int nb,
                  const float* m1, const float* m2, float* student_grads, float* loss_out,
                  void* stream);

/* =============================================================================================
 * Prioritized-replay sum-tree.  NO REFERENCE IMPLEMENTATION EXISTS (the reference only
 * cites Schaul et al. 2015 at README.md:15): parity is against oracle/sumtree.c, "parity unpinned".
 * Layout: implicit binary heap in one fp32 array tree[2*cap] (cap a power of two): node 1 is
 * the root, leaves are tree[cap + i]; parent = fp32(left + right), recomputed (never delta
 * updated) so results are bit-reproducible.
 * ============================================================================================= */
/* Set leaf idx[j] = prio[j] for j in [0,n) (last writer wins on duplicates: the highest j), then
 * recompute every affected ancestor level by level.  idx int64; entries outside [0,cap) are ignored.
 * n <= 2048: one CTA (in-kernel sort by leaf); larger batches: 128 CTAs, one per depth-7 subtree, + a 7-level top pass. */
int drl_sumtree_update(float* tree, int64_t cap, const int64_t* idx, const float* prio, int64_t n,
                       void* stream);
/* Rebuild all internal nodes from the leaves. */
int drl_sumtree_rebuild(float* tree, int64_t cap, void* stream);
/* Prefix-sum descent for n query masses u[j] in [0, total): at each node go left if u < left
 * else subtract left and go right.  idx_out[j] = leaf index, prio_out[j] = its priority. */
int drl_sumtree_sample(const float* tree, int64_t cap, const float* u, int64_t n,
                       int64_t* idx_out, float* prio_out, void* stream);
/* Stratified masses (Schaul et al. 2015 §3.3): u[j] = (j + uniform01[j]) * total / n, computed
 * on device in fp32 from caller-supplied uniforms so the host RNG stays the seed authority. */
int drl_sumtree_stratified_u(const float* tree, const float* uniform01, int64_t n, float* u_out,
                             void* stream);
/* One batch (n <= 1024) of proportional prioritized sampling in one launch (Schaul et al. 2015, Algorithm 1 lines 9-10):
 * stratified masses from uniform01, prefix-sum descent, importance weights w_j = (size * p_j / total)^-beta / max_k w_k
 * (0 for an empty tree / zero-priority leaf).  idx_out / prio_out bit-identical to stratified_u + sample. */
int drl_per_sample(const float* tree, int64_t cap, const float* uniform01, int64_t n, float size, float beta,
                   int64_t* idx_out, float* prio_out, float* weights_out, void* stream);
/* out[i] = ring[(idx[i] + shift) mod cap] for rows of row_bytes (multiple of 16, 16-byte aligned buffers): the sampled
 * transitions of a device-resident replay ring (uint8 frame stacks: shift 0 = observation, 1 = successor); n <= 65535. */
int drl_ring_gather(const void* ring, int64_t cap, int64_t row_bytes, const int64_t* idx, int64_t n, int shift, void* out,
                    void* stream);
/* Measurement aid: cycles per dependent load when one thread chases `hops` links through `ring` (int32 next-slot indices of a
 * cyclic permutation prepared by the caller; its size decides whether the chain lives in L2 or HBM).
 * cycles_per_hop_out[0] (device) receives the figure: levels x this is the latency floor of the sum-tree walks. */
int drl_probe_dependent_latency(const int* ring, int hops, float* cycles_per_hop_out, void* stream);

/* =============================================================================================
 * Actor side — the steps immediately before the learner path, for all environments of a rank per call.
 *
 * drl_sample_actions_f32 / drl_net_act replace RolloutManager._choose_action (drl/algos/common/rollout.py:280-288: one
 * batch-1 forward + Categorical.sample() per environment step).  Categorical.sample() is torch.multinomial(probs, 1), i.e.
 * argmax_a probs[a] / q[a] with q ~ Exp(1): `noise` [n, A] supplies the q variates (then the draw equals the reference's draw
 * for the same generator state); noise == NULL generates them in the kernel with Philox4x32-10 keyed by `seed`, counter =
 * (row, `counter`) — pass the environment step number.  logp_out (nullable) receives log pi(a|s).
 * drl_net_act = drl_net_forward on rows `row_idx` of the frame store + the draw; logits_out / values_out are nullable.
 * --------------------------------------------------------------------------------------------- */
int drl_sample_actions_f32(const float* logits, int64_t n, int num_actions, const float* noise, uint64_t seed,
                           uint64_t counter, int64_t* actions_out, float* logp_out, void* stream);
int drl_net_act(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, const float* noise, uint64_t seed,
                uint64_t counter, int64_t* actions_out, float* logp_out, float* logits_out, float* values_out, void* stream);
/* Rollout.record / record_nexts (rollout.py:159-195) for all n_envs environments of a rank: writes time step `t` of the
 * device-resident stores laid out [n_envs, pitch_steps, ...] (uint8 frames 84x84x4, int64 actions, float dones and one float
 * array per reward stream).  Every source pointer is nullable (record_nexts passes frames + actions only) and may be device
 * memory or pinned host memory; obs_on_host != 0 copies the frames with one strided host-to-device copy straight into place. */
int drl_rollout_record(uint8_t* obs_store, int64_t* act_store, float* done_store, float* const* reward_stores, int n_streams,
                       int n_envs, int64_t pitch_steps, int t, const uint8_t* obs_t, const int64_t* a_t, const float* d_t,
                       const float* const* r_t, int obs_on_host, void* stream);
/* FrameStackWrapper (drl/envs/wrappers/stateless/frame_stack.py:50-59, num_stack = 4) on the device: slot t_dst of every
 * environment = slot t_src shifted by one channel + the newest single frame (newest [n_envs, 84, 84] uint8, device or pinned
 * host memory) as channel 3; reset_mask[e] (nullable): 1 = four copies of the newest frame (the environment was reset),
 * 2 = leave environment e untouched, 0 = shift + append.  t_src == t_dst is allowed (in place).
 * The environments hand over 7 KB per step instead of the 28 KB stack. */
int drl_rollout_stack_frames(uint8_t* obs_store, int n_envs, int64_t pitch_steps, int t_src, int t_dst, const uint8_t* newest,
                             const uint8_t* reset_mask, void* stream);
/* The extra_steps carry-over of Rollout.report (rollout.py:213-223): `count` time steps [src_t0, src_t0+count) of every
 * environment of one [n_envs, pitch, step_bytes] array to steps [dst_t0, ...) of another (step_bytes a multiple of 4). */
int drl_rollout_copy_steps(void* dst, const void* src, int n_envs, int64_t dst_pitch_steps, int64_t src_pitch_steps,
                           int64_t step_bytes, int dst_t0, int src_t0, int count, void* stream);

/* =============================================================================================
 * Gradient exchange — replaces the DistributedDataParallel wrapper (drl/utils/launch.py:310;
 * allreduce inside backward, drl/algos/ppo.py:233) and global_mean
 * (drl/algos/common/metrics.py:8-26).  One process per GPU; the communicator is created from
 * an NCCL unique id the caller broadcasts (e.g. through torch.distributed).
 * ============================================================================================= */
typedef struct drl_comm drl_comm_t;
#define DRL_NCCL_UNIQUE_ID_BYTES 128
int drl_comm_unique_id(uint8_t* id_host /* [128] */);
int drl_comm_create(drl_comm_t** out, const uint8_t* id_host, int rank, int world, int device);
int drl_comm_destroy(drl_comm_t* comm);
/* All-reduce over NVLink PEER MEMORY instead of NCCL (one node, <= 8 ranks): every rank exports an exchange buffer
 * (cudaIpc handle, 64 bytes) sized for `max_elems` floats, the caller all-gathers the handles (e.g. through
 * torch.distributed) and hands all of them to every rank; from then on every gradient all-reduce of at most
 * max_elems floats issued through this communicator (drl_net_set_comm, drl_grad_allreduce_bucketed) runs as a
 * two-shot reduce-scatter / all-gather in three small kernels: each GPU reads its 1/N slice from all peers,
 * writes the sums back to all of them.  The result is bit-identical on every rank.  All ranks must call
 * export, exchange, import, and then synchronise (a barrier) before the first collective. */
#define DRL_IPC_HANDLE_BYTES 64
#define DRL_MAX_P2P_RANKS 8
int drl_comm_p2p_export(drl_comm_t* comm, int64_t max_elems, uint8_t* handle_out /* [64] */);
int drl_comm_p2p_import(drl_comm_t* comm, const uint8_t* handles_all /* [world][64], rank order; NULL = back to NCCL */);
/* In-place SUM allreduce of a flat fp32 buffer in `n_buckets` buckets given by bucket_offsets
 * [n_buckets+1] (host, elements), issued in order on `stream`. */
/* Vector-Jacobian product of the network (trunk + heads): grads[p] = d( sum(dlogits * logits) + sum(dvalues * values) ) / dp
 * for the minibatch (obs, row_idx), dlogits [nb, A], dvalues [nb, K] fp32 on the device.  For callers that put their
 * own loss on top of drl_net_forward — what autograd does behind the reference's Agent.forward (agent.py:53-75).
 * Activations are recomputed; `grads` is overwritten. */
int drl_net_backward(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, const float* dlogits,
                     const float* dvalues, float* grads, void* stream);
/* The same for a head that lives outside the handle (the DQN action-value head): trunk gradients from dfeat [nb, 512] =
 * d loss / d(the features drl_net_features returned).  Activations are recomputed unless reuse_forward != 0, which asserts that
 * the LAST call on this handle was drl_net_features on exactly (obs, row_idx, nb); `grads` (flat, drl_net_param_count floats)
 * is overwritten, the slices of the handle's own policy / value heads stay zero. */
int drl_net_backward_features(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, const float* dfeat,
                              float* grads, int reuse_forward, void* stream);
/* Attach (or detach, comm = NULL) a communicator to a network handle: drl_net_ppo_step then returns gradients
 * that are already SUMMED over the ranks — what DistributedDataParallel's backward hooks do in the reference
 * (drl/utils/launch.py:310, drl/algos/ppo.py:233) — and overlaps the exchange with the backward pass: the
 * fc + heads range (95 % of the parameters) is all-reduced on a side stream as soon as its gradients are final
 * (after the fc weight-gradient kernel), while the convolution gradients are still being computed; the small
 * convolution range follows after the last kernel.  The caller's stream waits for both before returning work. */
int drl_net_set_comm(drl_net_t* net, drl_comm_t* comm);
int drl_grad_allreduce_bucketed(drl_comm_t* comm, float* flat, const int64_t* bucket_offsets_host,
                                int n_buckets, void* stream);

#ifdef __cplusplus
}
#endif
#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#endif /* DRL_B200_H_ */
