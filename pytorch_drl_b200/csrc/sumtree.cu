// Prioritized-replay sum-tree (Schaul et al. 2015, proportional variant; cited by the reference at
// README.md:15 but NOT implemented there — SURVEY F2; parity is against oracle/sumtree.c).
//
// tree[2*cap] fp32, implicit heap: root = tree[1], leaves = tree[cap + i].  Every parent is
// recomputed as fp32(left + right) from its children (never delta-updated), so the tree is a
// pure function of the leaves and results are bit-reproducible on any schedule.
//
// update : a batch is sorted by leaf inside the CTA (bitonic sort of (leaf, j) keys in shared memory), the highest j of each
//          leaf writes it ("last writer wins", like sequential in-order writes), then all updated leaves walk up the levels in
//          lock-step (duplicate parents recompute the same value).  Batches of up to 2048 run in ONE CTA (20 dependent L2 round
//          trips: latency bound, not an HBM kernel).  Larger batches run on 128 CTAs, one per depth-7 subtree: every CTA scans the
//          batch, keeps its subtree's entries in arrival order (flushing whenever 2048 are pending, so later entries still win),
//          walks them up to its subtree root; a 7-level kernel then rebuilds the 127 top nodes.
//          Out-of-range leaves (idx < 0 or >= cap) are ignored.
// sample : one thread per query, prefix-sum descent; the top levels are shared by all queries
//          and stay in L1/L2.  drl_per_sample fuses stratified masses, descent and importance weights for one batch.
#include "common.cuh"

namespace drl {

constexpr int kSumtreeThreads = 1024;
constexpr int kSumtreeSlots = 2048;             // pending (leaf, j) keys of one flush
constexpr int kSumtreeSubBits = 7;              // multi-CTA update: 128 subtrees
constexpr uint64_t kKeyPad = ~uint64_t(0);

// Sorts keys[0 .. m) (m a power of two <= 2048, padded with kKeyPad) ascending; 1024 threads.
__device__ __forceinline__ void bitonic_sort_keys(uint64_t* keys, int m) {
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < m; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const uint64_t a = keys[t], b = keys[p];
          const bool up = (t & k) == 0;
          if ((a > b) == up) { keys[t] = b; keys[p] = a; }
        }
      }
      __syncthreads();
    }
  }
}

// keys[0 .. count): (leaf << 32) | j.  Writes each leaf once (highest j) and recomputes `levels` levels of ancestors.
__device__ __forceinline__ void sumtree_flush(float* __restrict__ tree, int64_t cap, const float* __restrict__ prio,
                                              uint64_t* keys, int count, int levels) {
  int m = 2;
  while (m < count) m <<= 1;
  for (int t = count + threadIdx.x; t < m; t += blockDim.x) keys[t] = kKeyPad;
  __syncthreads();
  bitonic_sort_keys(keys, m);
  int64_t node[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int t = threadIdx.x + q * kSumtreeThreads;
    node[q] = -1;
    if (t < count) {
      const uint64_t k = keys[t];
      const bool winner = t + 1 >= count || (keys[t + 1] >> 32) != (k >> 32);       // last (= highest j) of its leaf
      if (winner) {
        node[q] = cap + (int64_t)(k >> 32);
        tree[node[q]] = prio[(uint32_t)k];
      }
    }
  }
  __syncthreads();
  for (int l = 0; l < levels; ++l) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (node[q] >= 0) {
        node[q] >>= 1;
        // all children of this level were finalised before the barrier below (previous level)
        tree[node[q]] = __fadd_rn(tree[2 * node[q]], tree[2 * node[q] + 1]);
      }
    }
    __syncthreads();
  }
}

// n <= kSumtreeSlots entries, one CTA, all levels
__global__ void __launch_bounds__(kSumtreeThreads) sumtree_update_kernel(
    float* __restrict__ tree, int64_t cap, const int64_t* __restrict__ idx,
    const float* __restrict__ prio, int n, int levels) {
  __shared__ uint64_t keys[kSumtreeSlots];
  __shared__ int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int64_t leaf = idx[j];
    if (leaf >= 0 && leaf < cap) keys[atomicAdd(&s_count, 1)] = ((uint64_t)leaf << 32) | (uint32_t)j;   // order restored by the sort
  }
  __syncthreads();
  const int count = s_count;
  if (count > 0) sumtree_flush(tree, cap, prio, keys, count, levels);
}

// any n: CTA c owns the leaves [c << sub_levels, (c + 1) << sub_levels) and the subtree above them
__global__ void __launch_bounds__(kSumtreeThreads) sumtree_update_subtree_kernel(
    float* __restrict__ tree, int64_t cap, const int64_t* __restrict__ idx,
    const float* __restrict__ prio, int64_t n, int sub_levels) {
  __shared__ uint64_t keys[kSumtreeSlots];
  __shared__ int s_count;
  const int64_t c = blockIdx.x;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += kSumtreeThreads) {
    const int64_t j = base + threadIdx.x;
    const int64_t leaf = j < n ? idx[j] : -1;
    const bool mine = leaf >= 0 && leaf < cap && (leaf >> sub_levels) == c;
    if (__syncthreads_count(mine) == 0) continue;
    if (mine) keys[atomicAdd(&s_count, 1)] = ((uint64_t)leaf << 32) | (uint32_t)j;      // at most 1024 more: the buffer holds 2048
    __syncthreads();
    if (s_count > kSumtreeSlots - kSumtreeThreads) {           // the next tile might not fit: apply what is pending (earlier j first)
      sumtree_flush(tree, cap, prio, keys, s_count, sub_levels);
      if (threadIdx.x == 0) s_count = 0;
      __syncthreads();
    }
  }
  if (s_count > 0) sumtree_flush(tree, cap, prio, keys, s_count, sub_levels);
}

// the 2^bits - 1 nodes above the subtree roots, bottom-up
__global__ void sumtree_top_kernel(float* __restrict__ tree, int bits) {
  for (int l = bits - 1; l >= 0; --l) {
    const int node = (1 << l) + threadIdx.x;
    if ((int)threadIdx.x < (1 << l)) tree[node] = __fadd_rn(tree[2 * node], tree[2 * node + 1]);
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

// One batch of proportional prioritized sampling (Schaul et al. 2015, Algorithm 1 lines 9-10) in ONE CTA: stratified masses
// u_j = (j + uniform_j) * total / n, prefix-sum descent, P(j) = p_j / total, w_j = (size * P(j))^-beta / max_k w_k.
// An empty tree (total == 0) or a zero-priority leaf gives weight 0 instead of inf / NaN.
__global__ void __launch_bounds__(1024) per_sample_kernel(const float* __restrict__ tree, int64_t cap,
                                                         const float* __restrict__ uniform01, int n, float size, float beta,
                                                         int64_t* __restrict__ idx_out, float* __restrict__ prio_out,
                                                         float* __restrict__ w_out) {
  __shared__ float s_max[32];
  const int j = threadIdx.x;
  const float total = tree[1];
  float w = 0.f;
  if (j < n) {
    const float seg = __fdiv_rn(total, (float)n);
    float x = __fmul_rn(__fadd_rn((float)j, uniform01[j]), seg);
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
    const float p = __ldg(tree + node);
    idx_out[j] = node - cap;
    prio_out[j] = p;
    if (total > 0.f && p > 0.f) w = powf(__fmul_rn(size, __fdiv_rn(p, total)), -beta);
  }
  float m = warp_max(w);
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = warp_max(threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0.f);
    if (threadIdx.x == 0) s_max[0] = m;
  }
  __syncthreads();
  if (j < n) w_out[j] = s_max[0] > 0.f ? __fdiv_rn(w, s_max[0]) : 0.f;
}

// dst[i] = src[(idx[i] + shift) mod cap] for rows of row_bytes (multiple of 16): the sampled transitions of a replay ring
// (shift 1 = the successor frame stack)
__global__ void __launch_bounds__(256) ring_gather_kernel(const uint4* __restrict__ src, int64_t cap, int64_t row_vec,
                                                          const int64_t* __restrict__ idx, int shift, uint4* __restrict__ dst) {
  const int64_t i = blockIdx.y;
  int64_t r = (idx[i] + shift) % cap;
  if (r < 0) r += cap;
  const uint4* s = src + r * row_vec;
  uint4* d = dst + i * row_vec;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < row_vec; v += (int64_t)gridDim.x * blockDim.x) d[v] = __ldg(s + v);
}

// Dependent-access latency probe: one thread chases `hops` pointers through a ring of `n` int32 slots (a random cyclic permutation
// prepared by the caller) and reports cycles per hop — the per-level cost floor of the sum-tree walks (bench.py prints
// levels x this next to the measured update / sample times).
__global__ void latency_probe_kernel(const int* __restrict__ ring, int hops, float* __restrict__ cycles_per_hop) {
  int p = 0;
  for (int i = 0; i < 64; ++i) p = ring[p];                      // warm the path
  const long long t0 = clock64();
  for (int i = 0; i < hops; ++i) p = ring[p];
  const long long t1 = clock64();
  cycles_per_hop[0] = (float)(t1 - t0) / (float)hops;
  cycles_per_hop[1] = (float)p;                                  // keeps the chain alive
}

static int ilog2_exact(int64_t cap) {
  if (cap < 1 || (cap & (cap - 1))) return -1;
  int l = 0;
  while ((int64_t(1) << l) < cap) ++l;
  return l;
}

}  // namespace drl

using namespace drl;
namespace drl { extern int g_debug_flags[16]; }

extern "C" int drl_sumtree_update(float* tree, int64_t cap, const int64_t* idx, const float* prio,
                                  int64_t n, void* stream) {
  if (!tree || (n > 0 && (!idx || !prio))) return DRL_ERR_BAD_ARG;
  const int levels = ilog2_exact(cap);
  if (levels < 0 || n < 0 || n > (int64_t(1) << 31)) return DRL_ERR_BAD_SHAPE;
  if (n == 0) return DRL_OK;
  cudaStream_t st = as_stream(stream);
  // one CTA up to g_debug_flags[10] entries (default 512: measured 19.8 us at 512 / 46 us at 2048 on one CTA vs ~16 us on 128 CTAs)
  const int64_t single_max = g_debug_flags[10] > 0 ? (g_debug_flags[10] < kSumtreeSlots ? g_debug_flags[10] : kSumtreeSlots) : 512;
  if (n <= single_max || levels <= kSumtreeSubBits + 1) {
    // chunks are stream-ordered, so a later chunk overwrites an earlier one: global "highest j wins"
    for (int64_t off = 0; off < n; off += kSumtreeSlots) {
      const int m = (int)((n - off) < kSumtreeSlots ? (n - off) : kSumtreeSlots);
      sumtree_update_kernel<<<1, kSumtreeThreads, 0, st>>>(tree, cap, idx + off, prio + off, m, levels);
      DRL_LAUNCH_CHECK();
    }
    return DRL_OK;
  }
  sumtree_update_subtree_kernel<<<1 << kSumtreeSubBits, kSumtreeThreads, 0, st>>>(tree, cap, idx, prio, n, levels - kSumtreeSubBits);
  DRL_LAUNCH_CHECK();
  sumtree_top_kernel<<<1, 1 << (kSumtreeSubBits - 1), 0, st>>>(tree, kSumtreeSubBits);
  DRL_LAUNCH_CHECK();
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

extern "C" int drl_per_sample(const float* tree, int64_t cap, const float* uniform01, int64_t n, float size, float beta,
                              int64_t* idx_out, float* prio_out, float* weights_out, void* stream) {
  if (!tree || !uniform01 || !idx_out || !prio_out || !weights_out) return DRL_ERR_BAD_ARG;
  if (ilog2_exact(cap) < 0 || n < 1 || n > 1024) return DRL_ERR_BAD_SHAPE;
  per_sample_kernel<<<1, (unsigned)((n + 31) / 32 * 32), 0, as_stream(stream)>>>(tree, cap, uniform01, (int)n, size, beta, idx_out,
                                                                            prio_out, weights_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_ring_gather(const void* ring, int64_t cap, int64_t row_bytes, const int64_t* idx, int64_t n, int shift,
                               void* out, void* stream) {
  if (!ring || !idx || !out) return DRL_ERR_BAD_ARG;
  if (cap < 1 || n < 0 || row_bytes < 16 || n > 65535) return DRL_ERR_BAD_SHAPE;
  if ((row_bytes & 15) || (reinterpret_cast<uintptr_t>(ring) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return DRL_ERR_MISALIGNED;
  if (n == 0) return DRL_OK;
  const int64_t row_vec = row_bytes / 16;
  int gx = (int)((row_vec + 255) / 256);
  if (gx > 8) gx = 8;
  ring_gather_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint4*>(ring), cap, row_vec, idx,
                                                                                 shift, reinterpret_cast<uint4*>(out));
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_probe_dependent_latency(const int* ring, int hops, float* cycles_per_hop_out, void* stream) {
  if (!ring || !cycles_per_hop_out || hops < 1) return DRL_ERR_BAD_ARG;
  latency_probe_kernel<<<1, 1, 0, as_stream(stream)>>>(ring, hops, cycles_per_hop_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
