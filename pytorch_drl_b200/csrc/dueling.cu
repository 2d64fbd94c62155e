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
constexpr int kDuelMaxIn = 512, kDuelMaxP = 64, kDuelMaxH = 32;

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
  static size_t configured = 0;
  if (configured < smem) {
    DRL_CUDA(cudaFuncSetAttribute(dueling_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
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
