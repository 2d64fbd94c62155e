// Prioritized-replay sum-tree (Schaul et al. 2015, proportional variant; cited by the reference at
// README.md:15 but NOT implemented there — SURVEY F2; parity is against oracle/sumtree.c).
//
// tree[2*cap] fp32, implicit heap: root = tree[1], leaves = tree[cap + i].  Every parent is
// recomputed as fp32(left + right) from its children (never delta-updated), so the tree is a
// pure function of the leaves and results are bit-reproducible on any schedule.
//
// update : one CTA walks a batch of up to 1024 leaves up the 20 levels in lock-step
//          (__syncthreads between levels; duplicate parents recompute the same value).
//          Dependent-access latency bound: ~log2(cap) L2 round trips, not an HBM kernel.
// sample : one thread per query, prefix-sum descent; the top levels are shared by all queries
//          and stay in L1/L2.
#include "common.cuh"

namespace drl {

constexpr int kSumtreeChunk = 1024;

__global__ void __launch_bounds__(kSumtreeChunk) sumtree_update_kernel(
    float* __restrict__ tree, int64_t cap, const int64_t* __restrict__ idx,
    const float* __restrict__ prio, int n, int levels) {
  __shared__ int64_t s_idx[kSumtreeChunk];
  const int j = threadIdx.x;
  int64_t my = -1;
  if (j < n) { my = idx[j]; s_idx[j] = my; }
  __syncthreads();
  // duplicates: the highest j wins (same rule as sequential in-order writes)
  bool winner = j < n;
  if (winner) {
    for (int q = j + 1; q < n; ++q) if (s_idx[q] == my) { winner = false; break; }
  }
  if (winner) tree[cap + my] = prio[j];
  __syncthreads();
  int64_t node = cap + my;
  for (int l = 0; l < levels; ++l) {
    if (j < n) {
      node >>= 1;
      // all children of this level were finalised before the barrier below (previous level)
      const float s = __fadd_rn(tree[2 * node], tree[2 * node + 1]);
      tree[node] = s;
    }
    __syncthreads();
  }
}

__global__ void sumtree_level_kernel(float* __restrict__ tree, int64_t first, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) {
    const int64_t node = first + i;
    tree[node] = __fadd_rn(tree[2 * node], tree[2 * node + 1]);
  }
}

__global__ void sumtree_sample_kernel(const float* __restrict__ tree, int64_t cap,
                                      const float* __restrict__ u, int64_t n,
                                      int64_t* __restrict__ idx_out, float* __restrict__ prio_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float x = u[j];
  int64_t node = 1;
  while (node < cap) {
    const float left = __ldg(tree + 2 * node);
    const float right = __ldg(tree + 2 * node + 1);
    if (x < left || !(right > 0.f)) {
      node = 2 * node;
    } else {
      x = __fsub_rn(x, left);
      node = 2 * node + 1;
    }
  }
  idx_out[j] = node - cap;
  prio_out[j] = __ldg(tree + node);
}

__global__ void sumtree_stratified_kernel(const float* __restrict__ tree,
                                          const float* __restrict__ uniform01, int64_t n,
                                          float* __restrict__ u_out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float seg = __fdiv_rn(tree[1], (float)n);
  u_out[j] = __fmul_rn(__fadd_rn((float)j, uniform01[j]), seg);
}

static int ilog2_exact(int64_t cap) {
  if (cap < 1 || (cap & (cap - 1))) return -1;
  int l = 0;
  while ((int64_t(1) << l) < cap) ++l;
  return l;
}

}  // namespace drl

using namespace drl;

extern "C" int drl_sumtree_update(float* tree, int64_t cap, const int64_t* idx, const float* prio,
                                  int64_t n, void* stream) {
  if (!tree || (n > 0 && (!idx || !prio))) return DRL_ERR_BAD_ARG;
  const int levels = ilog2_exact(cap);
  if (levels < 0 || n < 0) return DRL_ERR_BAD_SHAPE;
  // chunks are stream-ordered, so a later chunk overwrites an earlier one: global "highest j wins"
  for (int64_t off = 0; off < n; off += kSumtreeChunk) {
    const int m = (int)((n - off) < kSumtreeChunk ? (n - off) : kSumtreeChunk);
    sumtree_update_kernel<<<1, kSumtreeChunk, 0, as_stream(stream)>>>(tree, cap, idx + off,
                                                                    prio + off, m, levels);
    DRL_LAUNCH_CHECK();
  }
  return DRL_OK;
}

extern "C" int drl_sumtree_rebuild(float* tree, int64_t cap, void* stream) {
  if (!tree) return DRL_ERR_BAD_ARG;
  if (ilog2_exact(cap) < 0) return DRL_ERR_BAD_SHAPE;
  for (int64_t count = cap / 2; count >= 1; count /= 2) {
    const int threads = 256;
    const int64_t blocks = (count + threads - 1) / threads;
    sumtree_level_kernel<<<(unsigned)blocks, threads, 0, as_stream(stream)>>>(tree, count, count);
    DRL_LAUNCH_CHECK();
  }
  return DRL_OK;
}

extern "C" int drl_sumtree_sample(const float* tree, int64_t cap, const float* u, int64_t n,
                                  int64_t* idx_out, float* prio_out, void* stream) {
  if (!tree || (n > 0 && (!u || !idx_out || !prio_out))) return DRL_ERR_BAD_ARG;
  if (ilog2_exact(cap) < 0 || n < 0) return DRL_ERR_BAD_SHAPE;
  if (n == 0) return DRL_OK;
  const int threads = 128;
  sumtree_sample_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, as_stream(stream)>>>(
      tree, cap, u, n, idx_out, prio_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_sumtree_stratified_u(const float* tree, const float* uniform01, int64_t n,
                                        float* u_out, void* stream) {
  if (!tree || (n > 0 && (!uniform01 || !u_out))) return DRL_ERR_BAD_ARG;
  if (n < 0) return DRL_ERR_BAD_SHAPE;
  if (n == 0) return DRL_OK;
  const int threads = 128;
  sumtree_stratified_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0,
                              as_stream(stream)>>>(tree, uniform01, n, u_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
