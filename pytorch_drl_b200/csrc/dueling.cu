// DQN head blocks the reference ships: the dueling action-value head and the epsilon-greedy policy
// built on it, for a whole batch in one launch each.
//
// drl_dueling_forward_f32 replaces DuelingArchitecture.forward (drl/agents/architectures/stateless/dueling.py:50-57)
// behind SimpleDiscreteActionValueHead.forward (drl/agents/heads/action_value_heads.py:124-140):
//   proj = relu(W0 f + b0); adv = MLP_a(proj); adv -= mean(adv); val = MLP_v(proj); q = adv + val
// where MLP (mlp.py:32-41, num_layers = 2) is Linear -> ReLU -> Linear -> ReLU — both streams end in a ReLU.
// One warp per sample; all weights live in shared memory (the 512 x 50 projection transposed so that lane p reads
// column p without bank conflicts), the 512 features are read once.  fp32 CUDA cores: 28 k MACs per sample.
//
// drl_epsilon_greedy_probs_f32 replaces the arithmetic of EpsilonGreedyCategoricalPolicyHead.forward
// (drl/agents/heads/policy_heads.py:219-224): (1 - eps) * one_hot(argmax q) + eps / A, first maximum on ties.
#include "common.cuh"

namespace drl {

constexpr int kDuelWarps = 8;
constexpr int kDuelMaxIn = 512;

struct DuelingParams {
  const float *w0, *b0, *wa1, *ba1, *wa2, *ba2, *wv1, *bv1, *wv2, *bv2;
};

__global__ void __launch_bounds__(kDuelWarps * 32) dueling_forward_kernel(const float* __restrict__ feat, int nb, int in_dim, int P,
                                                                          int H, int A, DuelingParams p, float* __restrict__ q_out) {
  extern __shared__ float sm[];
  float* w0t = sm;                          // [in_dim][P]
  float* wa1t = w0t + in_dim * P;           // [P][H]
  float* wv1t = wa1t + P * H;               // [P][H]
  float* wa2 = wv1t + P * H;                // [A][H]
  float* wv2 = wa2 + A * H;                 // [H]
  float* bias = wv2 + H;                    // b0[P] ba1[H] ba2[A] bv1[H] bv2[1]
  float* scratch = bias + P + 2 * H + A + 1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sf = scratch + warp * (in_dim + P + 2 * H);   // features, proj, adv hidden, val hidden of this warp's sample
  float* sproj = sf + in_dim;
  float* sah = sproj + P;
  float* svh = sah + H;

  for (int e = threadIdx.x; e < in_dim * P; e += blockDim.x) { const int pp = e / in_dim, k = e - pp * in_dim; w0t[k * P + pp] = p.w0[e]; }
  for (int e = threadIdx.x; e < P * H; e += blockDim.x) {
    const int h = e / P, pp = e - h * P;
    wa1t[pp * H + h] = p.wa1[e];
    wv1t[pp * H + h] = p.wv1[e];
  }
  for (int e = threadIdx.x; e < A * H; e += blockDim.x) wa2[e] = p.wa2[e];
  for (int e = threadIdx.x; e < H; e += blockDim.x) { wv2[e] = p.wv2[e]; bias[P + e] = p.ba1[e]; bias[P + H + A + e] = p.bv1[e]; }
  for (int e = threadIdx.x; e < P; e += blockDim.x) bias[e] = p.b0[e];
  for (int e = threadIdx.x; e < A; e += blockDim.x) bias[P + H + e] = p.ba2[e];
  if (threadIdx.x == 0) bias[P + 2 * H + A] = p.bv2[0];
  __syncthreads();

  const int p0 = lane, p1 = lane + 32;
  for (int i = blockIdx.x * kDuelWarps + warp; i < nb; i += gridDim.x * kDuelWarps) {
    for (int k = lane; k < in_dim; k += 32) sf[k] = feat[(size_t)i * in_dim + k];
    __syncwarp();
    float a0 = p0 < P ? bias[p0] : 0.f, a1 = p1 < P ? bias[p1] : 0.f;
    for (int k = 0; k < in_dim; ++k) {
      const float f = sf[k];
      if (p0 < P) a0 = fmaf(f, w0t[k * P + p0], a0);
      if (p1 < P) a1 = fmaf(f, w0t[k * P + p1], a1);
    }
    if (p0 < P) sproj[p0] = fmaxf(a0, 0.f);
    if (p1 < P) sproj[p1] = fmaxf(a1, 0.f);
    __syncwarp();
    if (lane < H) {
      float ah = bias[P + lane], vh = bias[P + H + A + lane];
      for (int pp = 0; pp < P; ++pp) {
        const float x = sproj[pp];
        ah = fmaf(x, wa1t[pp * H + lane], ah);
        vh = fmaf(x, wv1t[pp * H + lane], vh);
      }
      sah[lane] = fmaxf(ah, 0.f);
      svh[lane] = fmaxf(vh, 0.f);
    }
    __syncwarp();
    float adv = 0.f, val = bias[P + 2 * H + A];
    if (lane < A) {
      adv = bias[P + H + lane];
      for (int h = 0; h < H; ++h) adv = fmaf(sah[h], wa2[lane * H + h], adv);
      adv = fmaxf(adv, 0.f);                       // the MLP's trailing ReLU (mlp.py:32-41)
    }
    for (int h = 0; h < H; ++h) val = fmaf(svh[h], wv2[h], val);
    val = fmaxf(val, 0.f);
    const float mean = warp_sum(lane < A ? adv : 0.f) / (float)A;
    if (lane < A) q_out[(size_t)i * A + lane] = (adv - mean) + val;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(256) epsilon_greedy_kernel(const float* __restrict__ q, int nb, int A, float one_minus_eps, float eps,
                                                             float* __restrict__ probs, int64_t* __restrict__ greedy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  const float* qi = q + (size_t)i * A;
  int best = 0;
  float bv = qi[0];
  for (int a = 1; a < A; ++a) {
    const float v = qi[a];
    if (v > bv) { bv = v; best = a; }             // strict: first maximum, like torch.argmax
  }
  const float u = __fdiv_rn(1.f, (float)A);        // ones / ones.sum(-1)
  const float eu = __fmul_rn(eps, u);
  for (int a = 0; a < A; ++a) probs[(size_t)i * A + a] = a == best ? __fadd_rn(one_minus_eps, eu) : eu;
  if (greedy) greedy[i] = best;
}


// ---- DQN learner step on the dueling head (BASELINE config "Double/Dueling DQN with prioritized replay") ------------------
// The reference ships the blocks (dueling.py:9-57, action_value_heads.py:85-140, the double-Q Bellman operator
// credit_assignment.py:252-305) but no DQN algorithm: this is the loss a learner puts on them,
//     y_i     = r_i + (1 - d_i) * gamma * Q_tgt(s'_i, argmax_a Q_sel(s'_i, a))          Q_sel = online net (double-Q) or target net
//     delta_i = Q(s_i, a_i) - y_i
//     loss    = mean_i w_i * crit(delta_i)      crit = SmoothL1 (beta 1) or squared error (drl/algos/common/losses.py:6-18, 'none' reduction)
// and its gradient through the head down to the trunk features.  Kernel 1 (one warp per sample, weights in shared memory like the
// forward kernel) recomputes the head forward, forms delta / loss, backpropagates through the two 25-unit streams, accumulates
// their weight gradients in shared memory (one global atomic per element and CTA) and writes d(pre-activation of the projection)
// [nb, P].  Kernels 2 / 3 are the two thin GEMMs of the projection layer: gW0 = dproj^T feat, dfeat = dproj W0.
struct DuelingGrads {
  float *w0, *b0, *wa1, *ba1, *wa2, *ba2, *wv1, *bv1, *wv2, *bv2;
};

__global__ void __launch_bounds__(kDuelWarps * 32) dueling_loss_backward_kernel(
    const float* __restrict__ feat, int nb, int in_dim, int P, int H, int A, DuelingParams p, DuelingGrads g,
    const int64_t* __restrict__ actions, const float* __restrict__ rewards, const float* __restrict__ dones,
    const float* __restrict__ q_next_sel, const float* __restrict__ q_next_tgt, const float* __restrict__ is_weights, float gamma,
    int use_dones, int smooth_l1, float* __restrict__ q_out, float* __restrict__ td_out, float* __restrict__ y_out,
    float* __restrict__ loss_out, float* __restrict__ dproj_out) {
  extern __shared__ float sm[];
  float* w0t = sm;                          // [in_dim][P]
  float* wa1t = w0t + in_dim * P;           // [P][H]
  float* wv1t = wa1t + P * H;               // [P][H]
  float* wa2 = wv1t + P * H;                // [A][H]
  float* wv2 = wa2 + A * H;                 // [H]
  float* bias = wv2 + H;                    // b0[P] ba1[H] ba2[A] bv1[H] bv2[1]
  float* gacc = bias + P + 2 * H + A + 1;   // gradient accumulators: gwa1[H][P] gwv1[H][P] gwa2[A][H] gwv2[H] gba1[H] gba2[A] gbv1[H] gbv2[1]
  const int n_gacc = 2 * H * P + A * H + H + H + A + H + 1;
  float* scratch = gacc + n_gacc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sf = scratch + warp * (in_dim + 2 * P + 4 * H + 32);
  float* sproj = sf + in_dim;               // post-ReLU projection
  float* sah = sproj + P;                   // post-ReLU hidden units of the two streams
  float* svh = sah + H;
  float* dah = svh + H;                     // d(pre-activation) of the hidden units
  float* dvh = dah + H;
  float* sda = dvh + H;                     // [32] d(pre-activation) of the advantage outputs

  for (int e = threadIdx.x; e < in_dim * P; e += blockDim.x) { const int pp = e / in_dim, k = e - pp * in_dim; w0t[k * P + pp] = p.w0[e]; }
  for (int e = threadIdx.x; e < P * H; e += blockDim.x) {
    const int h = e / P, pp = e - h * P;
    wa1t[pp * H + h] = p.wa1[e];
    wv1t[pp * H + h] = p.wv1[e];
  }
  for (int e = threadIdx.x; e < A * H; e += blockDim.x) wa2[e] = p.wa2[e];
  for (int e = threadIdx.x; e < H; e += blockDim.x) { wv2[e] = p.wv2[e]; bias[P + e] = p.ba1[e]; bias[P + H + A + e] = p.bv1[e]; }
  for (int e = threadIdx.x; e < P; e += blockDim.x) bias[e] = p.b0[e];
  for (int e = threadIdx.x; e < A; e += blockDim.x) bias[P + H + e] = p.ba2[e];
  if (threadIdx.x == 0) bias[P + 2 * H + A] = p.bv2[0];
  for (int e = threadIdx.x; e < n_gacc; e += blockDim.x) gacc[e] = 0.f;
  __syncthreads();
  float* gwa1 = gacc;                       // [H][P] (torch layout of the gradient)
  float* gwv1 = gwa1 + H * P;
  float* gwa2 = gwv1 + H * P;               // [A][H]
  float* gwv2 = gwa2 + A * H;
  float* gba1 = gwv2 + H;
  float* gba2 = gba1 + H;
  float* gbv1 = gba2 + A;
  float* gbv2 = gbv1 + H;

  const int p0 = lane, p1 = lane + 32;
  const float inv_nb = 1.f / (float)nb;
  float loss_local = 0.f;
  for (int i = blockIdx.x * kDuelWarps + warp; i < nb; i += gridDim.x * kDuelWarps) {
    // ---- forward (as dueling_forward_kernel) ----
    for (int k = lane; k < in_dim; k += 32) sf[k] = feat[(size_t)i * in_dim + k];
    __syncwarp();
    float a0 = p0 < P ? bias[p0] : 0.f, a1 = p1 < P ? bias[p1] : 0.f;
    for (int k = 0; k < in_dim; ++k) {
      const float f = sf[k];
      if (p0 < P) a0 = fmaf(f, w0t[k * P + p0], a0);
      if (p1 < P) a1 = fmaf(f, w0t[k * P + p1], a1);
    }
    if (p0 < P) sproj[p0] = fmaxf(a0, 0.f);
    if (p1 < P) sproj[p1] = fmaxf(a1, 0.f);
    __syncwarp();
    if (lane < H) {
      float ah = bias[P + lane], vh = bias[P + H + A + lane];
      for (int pp = 0; pp < P; ++pp) {
        const float x = sproj[pp];
        ah = fmaf(x, wa1t[pp * H + lane], ah);
        vh = fmaf(x, wv1t[pp * H + lane], vh);
      }
      sah[lane] = fmaxf(ah, 0.f);
      svh[lane] = fmaxf(vh, 0.f);
    }
    __syncwarp();
    float adv = 0.f, val = bias[P + 2 * H + A];
    if (lane < A) {
      adv = bias[P + H + lane];
      for (int h = 0; h < H; ++h) adv = fmaf(sah[h], wa2[lane * H + h], adv);
      adv = fmaxf(adv, 0.f);
    }
    for (int h = 0; h < H; ++h) val = fmaf(svh[h], wv2[h], val);
    val = fmaxf(val, 0.f);
    const float mean = warp_sum(lane < A ? adv : 0.f) / (float)A;
    const float q = (adv - mean) + val;
    if (q_out && lane < A) q_out[(size_t)i * A + lane] = q;
    // ---- Bellman target (credit_assignment.py:283-303 with n = 1), TD error, loss ----
    const float* qs = q_next_sel + (size_t)i * A;
    const float* qt = q_next_tgt + (size_t)i * A;
    int abest = 0;
    float best = qs[0];
    for (int j = 1; j < A; ++j) { const float v = qs[j]; if (v > best) { best = v; abest = j; } }     // first maximum, like torch.argmax
    // r + (nonterminal * gamma) * Q with the reference's float32 roundings (no FMA contraction), like credit.cu's nstep_backup
    const float ntg = use_dones ? __fmul_rn(__fsub_rn(1.f, dones[i]), gamma) : gamma;
    const float y = __fadd_rn(rewards[i], __fmul_rn(ntg, qt[abest]));
    const int act = (int)actions[i];
    const float q_taken = __shfl_sync(0xffffffffu, q, act);
    const float delta = q_taken - y;
    const float w = is_weights ? is_weights[i] : 1.f;
    float crit, dcrit;
    if (smooth_l1) {                                     // torch.nn.SmoothL1Loss(beta = 1)
      const float ad = fabsf(delta);
      crit = ad < 1.f ? 0.5f * delta * delta : ad - 0.5f;
      dcrit = ad < 1.f ? delta : (delta > 0.f ? 1.f : -1.f);
    } else {                                             // torch.nn.MSELoss
      crit = delta * delta;
      dcrit = 2.f * delta;
    }
    if (lane == 0) {
      td_out[i] = delta;
      if (y_out) y_out[i] = y;
      loss_local += w * crit * inv_nb;
    }
    const float gq = w * dcrit * inv_nb;                 // d loss / d Q(s_i, a_i)
    // ---- backward: q = adv - mean(adv) + val ----
    float dadv = 0.f;
    if (lane < A) dadv = (adv > 0.f) ? gq * ((lane == act ? 1.f : 0.f) - 1.f / (float)A) : 0.f;       // through the trailing ReLU
    const float dval = val > 0.f ? gq : 0.f;
    sda[lane] = dadv;
    __syncwarp();
    if (lane < A) {
      atomicAdd(gba2 + lane, dadv);
      for (int h = 0; h < H; ++h) atomicAdd(gwa2 + lane * H + h, dadv * sah[h]);
    }
    if (lane == 0) atomicAdd(gbv2, dval);
    if (lane < H) {
      atomicAdd(gwv2 + lane, dval * svh[lane]);
      // d(hidden) = W2^T d(out), through the hidden ReLU
      float da = 0.f;
      for (int a = 0; a < A; ++a) da = fmaf(sda[a], wa2[a * H + lane], da);
      da = sah[lane] > 0.f ? da : 0.f;
      const float dv = svh[lane] > 0.f ? dval * wv2[lane] : 0.f;
      dah[lane] = da;
      dvh[lane] = dv;
      atomicAdd(gba1 + lane, da);
      atomicAdd(gbv1 + lane, dv);
    }
    __syncwarp();
    // gWa1[h][p] += dah[h] * proj[p]; d(proj)[p] = sum_h (dah[h] Wa1[h][p] + dvh[h] Wv1[h][p]) through the projection's ReLU
    for (int pp = lane; pp < P; pp += 32) {
      const float x = sproj[pp];
      float dp = 0.f;
      for (int h = 0; h < H; ++h) {
        const float da = dah[h], dv = dvh[h];
        atomicAdd(gwa1 + h * P + pp, da * x);
        atomicAdd(gwv1 + h * P + pp, dv * x);
        dp = fmaf(da, wa1t[pp * H + h], dp);
        dp = fmaf(dv, wv1t[pp * H + h], dp);
      }
      dp = x > 0.f ? dp : 0.f;
      dproj_out[(size_t)i * P + pp] = dp;
    }
    __syncwarp();
  }
  if (lane == 0 && loss_local != 0.f) atomicAdd(loss_out, loss_local);
  __syncthreads();
  for (int e = threadIdx.x; e < H * P; e += blockDim.x) { atomicAdd(g.wa1 + e, gwa1[e]); atomicAdd(g.wv1 + e, gwv1[e]); }
  for (int e = threadIdx.x; e < A * H; e += blockDim.x) atomicAdd(g.wa2 + e, gwa2[e]);
  for (int e = threadIdx.x; e < H; e += blockDim.x) { atomicAdd(g.wv2 + e, gwv2[e]); atomicAdd(g.ba1 + e, gba1[e]); atomicAdd(g.bv1 + e, gbv1[e]); }
  for (int e = threadIdx.x; e < A; e += blockDim.x) atomicAdd(g.ba2 + e, gba2[e]);
  if (threadIdx.x == 0) atomicAdd(g.bv2, gbv2[0]);
}

// gW0[p][k] += sum_i dproj[i][p] * feat[i][k]   (thread = k, block row = p; the batch is split over blockIdx.z)
__global__ void __launch_bounds__(256) dueling_proj_wgrad_kernel(const float* __restrict__ dproj, const float* __restrict__ feat, int nb,
                                                                 int in_dim, int P, int rows_per_split, float* __restrict__ gw0,
                                                                 float* __restrict__ gb0) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, pp = blockIdx.y;
  const int i0 = blockIdx.z * rows_per_split, i1 = min(nb, i0 + rows_per_split);
  float acc = 0.f, accb = 0.f;
  if (k < in_dim) {
    for (int i = i0; i < i1; ++i) {
      const float d = dproj[(size_t)i * P + pp];
      acc = fmaf(d, feat[(size_t)i * in_dim + k], acc);
      accb += d;
    }
    atomicAdd(gw0 + (size_t)pp * in_dim + k, acc);
    if (k == 0) atomicAdd(gb0 + pp, accb);
  }
}

// dfeat[i][k] = sum_p dproj[i][p] * W0[p][k]
__global__ void __launch_bounds__(256) dueling_proj_dgrad_kernel(const float* __restrict__ dproj, const float* __restrict__ w0, int nb,
                                                                 int in_dim, int P, float* __restrict__ dfeat) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
  if (k >= in_dim) return;
  float acc = 0.f;
  for (int pp = 0; pp < P; ++pp) acc = fmaf(dproj[(size_t)i * P + pp], w0[(size_t)pp * in_dim + k], acc);
  dfeat[(size_t)i * in_dim + k] = acc;
}

constexpr int kDuelMaxDevices = 64;
template <typename Kern>
static int duel_ensure_smem(Kern kern, size_t smem, size_t (&configured)[kDuelMaxDevices]) {
  int dev = 0;
  DRL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kDuelMaxDevices) return DRL_ERR_BAD_ARG;
  if (configured[dev] < smem) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev] = smem;
  }
  return DRL_OK;
}

}  // namespace drl

using namespace drl;

extern "C" int drl_dueling_forward_f32(const float* feat, int nb, int in_dim, int num_actions, int widening, const float* w0,
                                       const float* b0, const float* wa1, const float* ba1, const float* wa2, const float* ba2,
                                       const float* wv1, const float* bv1, const float* wv2, const float* bv2, float* q_out,
                                       void* stream) {
  if (!feat || !q_out || !w0 || !b0 || !wa1 || !ba1 || !wa2 || !ba2 || !wv1 || !bv1 || !wv2 || !bv2) return DRL_ERR_BAD_ARG;
  if (nb < 1 || in_dim < 1 || in_dim > kDuelMaxIn || num_actions < 1 || num_actions > DRL_MAX_ACTIONS) return DRL_ERR_BAD_SHAPE;
  if (widening != 1) return DRL_ERR_BAD_SHAPE;      // 50 / 25 internal features (Wang et al. 2016); wider heads do not fit the smem plan
  const int P = 50 * widening, H = 25 * widening, A = num_actions;
  const size_t smem = sizeof(float) * ((size_t)in_dim * P + 2 * P * H + A * H + H + (P + 2 * H + A + 1) +
                                       (size_t)kDuelWarps * (in_dim + P + 2 * H));
  static size_t configured[kDuelMaxDevices] = {0};       // the attribute is a per-device setting
  DRL_TRY(duel_ensure_smem(dueling_forward_kernel, smem, configured));
  int blocks = (nb + kDuelWarps - 1) / kDuelWarps;
  if (blocks > sm_count()) blocks = sm_count();
  DuelingParams p{w0, b0, wa1, ba1, wa2, ba2, wv1, bv1, wv2, bv2};
  dueling_forward_kernel<<<blocks, kDuelWarps * 32, smem, as_stream(stream)>>>(feat, nb, in_dim, P, H, A, p, q_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_epsilon_greedy_probs_f32(const float* q, int nb, int num_actions, double epsilon, float* probs_out,
                                            int64_t* greedy_out, void* stream) {
  if (!q || !probs_out) return DRL_ERR_BAD_ARG;
  if (nb < 1 || num_actions < 1) return DRL_ERR_BAD_SHAPE;
  if (!(epsilon >= 0.0 && epsilon <= 1.0)) return DRL_ERR_BAD_ARG;      // policy_heads.py:185-188
  // (1 - eps) is formed in double like the Python expression, then both scalars are rounded to the tensor's float32
  const float ome = (float)(1.0 - epsilon);
  epsilon_greedy_kernel<<<(nb + 255) / 256, 256, 0, as_stream(stream)>>>(q, nb, num_actions, ome, (float)epsilon, probs_out, greedy_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_dueling_loss_backward_f32(const float* feat, int nb, int in_dim, int num_actions, int widening,
                                             const float* head_params, const int64_t* actions, const float* rewards,
                                             const float* dones, const float* q_next_online, const float* q_next_target,
                                             const float* is_weights, float gamma, int use_dones, int double_q, int smooth_l1,
                                             float* q_out, float* td_out, float* y_out, float* loss_out, float* dfeat_out,
                                             float* head_grads, float* workspace, void* stream) {
  if (!feat || !head_params || !actions || !rewards || !dones || !q_next_target || (double_q && !q_next_online) || !td_out ||
      !loss_out || !dfeat_out || !head_grads || !workspace)
    return DRL_ERR_BAD_ARG;
  if (nb < 1 || in_dim < 1 || in_dim > kDuelMaxIn || num_actions < 1 || num_actions > DRL_MAX_ACTIONS) return DRL_ERR_BAD_SHAPE;
  if (widening != 1) return DRL_ERR_BAD_SHAPE;
  const int P = 50 * widening, H = 25 * widening, A = num_actions;
  // flat layout of head_params / head_grads: the reference module's parameters() order (oracle/dueling_oracle.py PARAM_KEYS)
  auto carve = [&](const float* base, DuelingParams& o) {
    const float* q = base;
    o.w0 = q; q += (size_t)P * in_dim; o.b0 = q; q += P; o.wa1 = q; q += H * P; o.ba1 = q; q += H; o.wa2 = q; q += A * H; o.ba2 = q; q += A;
    o.wv1 = q; q += H * P; o.bv1 = q; q += H; o.wv2 = q; q += H; o.bv2 = q;
  };
  DuelingParams p;
  carve(head_params, p);
  DuelingParams gp;
  carve(head_grads, gp);
  DuelingGrads g{const_cast<float*>(gp.w0), const_cast<float*>(gp.b0), const_cast<float*>(gp.wa1), const_cast<float*>(gp.ba1),
                 const_cast<float*>(gp.wa2), const_cast<float*>(gp.ba2), const_cast<float*>(gp.wv1), const_cast<float*>(gp.bv1),
                 const_cast<float*>(gp.wv2), const_cast<float*>(gp.bv2)};
  cudaStream_t st = as_stream(stream);
  const int n_gacc = 2 * H * P + A * H + H + H + A + H + 1;
  const size_t smem = sizeof(float) * ((size_t)in_dim * P + 2 * P * H + A * H + H + (P + 2 * H + A + 1) + n_gacc +
                                       (size_t)kDuelWarps * (in_dim + 2 * P + 4 * H + 32));
  static size_t configured[kDuelMaxDevices] = {0};
  DRL_TRY(duel_ensure_smem(dueling_loss_backward_kernel, smem, configured));
  int blocks = (nb + kDuelWarps - 1) / kDuelWarps;
  if (blocks > sm_count()) blocks = sm_count();
  dueling_loss_backward_kernel<<<blocks, kDuelWarps * 32, smem, st>>>(
      feat, nb, in_dim, P, H, A, p, g, actions, rewards, dones, double_q ? q_next_online : q_next_target, q_next_target, is_weights,
      gamma, use_dones, smooth_l1, q_out, td_out, y_out, loss_out, workspace);
  DRL_LAUNCH_CHECK();
  const int splits = nb >= 256 ? 8 : 1, rows = (nb + splits - 1) / splits;
  dueling_proj_wgrad_kernel<<<dim3((in_dim + 255) / 256, P, splits), 256, 0, st>>>(workspace, feat, nb, in_dim, P, rows, g.w0, g.b0);
  DRL_LAUNCH_CHECK();
  dueling_proj_dgrad_kernel<<<dim3((in_dim + 255) / 256, nb), 256, 0, st>>>(workspace, p.w0, nb, in_dim, P, dfeat_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
