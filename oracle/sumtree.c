/*
 * CPU oracle for the prioritized-replay sum-tree (plain C, strict fp32, no FMA contraction).
 *
 * TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
 *
 * PARITY UNPINNED: the reference (lucaslingle/pytorch_drl) cites Schaul et al. 2015,
 * "Prioritized Experience Replay" (README.md:15) but contains no sum-tree, replay buffer or DQN
 * learner (drl/algos/__init__.py:1-4 exports only PPO), so there is no reference
 * code, test or golden vector to pin this against.  The semantics below are the builder's
 * statement of the paper's proportional variant (§3.3 + Appendix B.2.1: sum-tree, stratified
 * sampling of k equal-mass segments) and are what `pytorch_drl_b200/csrc/sumtree.cu` must match
 * bit for bit:
 *   - tree[2*cap] floats, root tree[1], leaves tree[cap+i]; parent = (float)(left + right),
 *     always recomputed from the children;
 *   - update: writes applied in order j = 0..n-1 (so the highest j wins on duplicates);
 *   - sample: at each node go left iff (u < left || !(right > 0)); else u -= left, go right;
 *   - stratified masses: seg = total / (float)n;  u_j = ((float)j + uniform_j) * seg.
 *   - prioritized sampling of one batch (Schaul et al. 2015, Algorithm 1 lines 9-10): the stratified masses above, the
 *     descent, P(j) = p_j / total, w_j = powf(size * P(j), -beta) normalised by the batch maximum; 0 for an empty tree or
 *     a zero-priority leaf.  Indices / priorities are bit-exact contracts, the weights go through powf (libm vs CUDA: compared
 *     with a relative tolerance).
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>

void st_rebuild(float* tree, int64_t cap) {
  for (int64_t node = cap - 1; node >= 1; --node) {
    volatile float s = tree[2 * node] + tree[2 * node + 1];
    tree[node] = s;
  }
}

void st_update(float* tree, int64_t cap, const int64_t* idx, const float* prio, int64_t n) {
  for (int64_t j = 0; j < n; ++j) {
    int64_t node = cap + idx[j];
    tree[node] = prio[j];
    for (node >>= 1; node >= 1; node >>= 1) {
      volatile float s = tree[2 * node] + tree[2 * node + 1];
      tree[node] = s;
    }
  }
}

void st_sample(const float* tree, int64_t cap, const float* u, int64_t n, int64_t* idx_out,
               float* prio_out) {
  for (int64_t j = 0; j < n; ++j) {
    float x = u[j];
    int64_t node = 1;
    while (node < cap) {
      const float left = tree[2 * node], right = tree[2 * node + 1];
      if (x < left || !(right > 0.0f)) {
        node = 2 * node;
      } else {
        volatile float d = x - left;
        x = d;
        node = 2 * node + 1;
      }
    }
    idx_out[j] = node - cap;
    prio_out[j] = tree[node];
  }
}

void st_stratified_u(const float* tree, const float* uniform01, int64_t n, float* u_out) {
  volatile float seg = tree[1] / (float)n;
  for (int64_t j = 0; j < n; ++j) {
    volatile float a = (float)j + uniform01[j];
    volatile float b = a * seg;
    u_out[j] = b;
  }
}

void st_per_sample(const float* tree, int64_t cap, const float* uniform01, int64_t n, float size, float beta,
                   int64_t* idx_out, float* prio_out, float* w_out) {
  const float total = tree[1];
  volatile float seg = total / (float)n;
  float wmax = 0.0f;
  for (int64_t j = 0; j < n; ++j) {
    volatile float a = (float)j + uniform01[j];
    volatile float u = a * seg;
    float x = u;
    int64_t node = 1;
    while (node < cap) {
      const float left = tree[2 * node], right = tree[2 * node + 1];
      if (x < left || !(right > 0.0f)) {
        node = 2 * node;
      } else {
        volatile float d = x - left;
        x = d;
        node = 2 * node + 1;
      }
    }
    const float p = tree[node];
    idx_out[j] = node - cap;
    prio_out[j] = p;
    float w = 0.0f;
    if (total > 0.0f && p > 0.0f) {
      volatile float prob = p / total;
      volatile float arg = size * prob;
      w = powf(arg, -beta);
    }
    w_out[j] = w;
    if (w > wmax) wmax = w;
  }
  for (int64_t j = 0; j < n; ++j) {
    volatile float q = wmax > 0.0f ? w_out[j] / wmax : 0.0f;
    w_out[j] = q;
  }
}
