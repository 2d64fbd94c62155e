// Actor side of the learner path: batched policy sampling for all environments of a rank per call and
// the device-resident rollout ring the actors write into.
//
//   reference                                                  here
//   RolloutManager._choose_action  (rollout.py:280-288)        drl_net_act / drl_sample_actions_f32 (all N envs per launch)
//   Rollout.record / record_nexts  (rollout.py:159-195)        drl_rollout_record  (one time step of all envs into [n_envs, Tp, ...])
//   Rollout.report carry-over      (rollout.py:197-224)        drl_rollout_copy_steps
//
// Sampling semantics: `Categorical(logits).sample()` is `torch.multinomial(probs, 1)`, which draws q ~ Exp(1) per class and
// returns argmax(probs / q).  The kernel does exactly that; the caller either hands in the q variates (then the draw is the
// reference's draw for the same generator state) or lets the kernel generate them with Philox4x32-10.
#include "common.cuh"
#include "net.cuh"

namespace drl {

constexpr int kObsBytes = 84 * 84 * 4;       // one NHWC uint8 frame stack
constexpr int kMaxSampleActions = 64;

__device__ __forceinline__ uint32_t mulhilo(uint32_t a, uint32_t b, uint32_t& hi) {
  const uint64_t p = (uint64_t)a * b;
  hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}

// Philox4x32-10 (Salmon et al. 2011): counter (c0..c3), key (k0, k1) -> 4 x 32 random bits
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo(0xD2511F53u, c.x, hi0);
    const uint32_t lo1 = mulhilo(0xCD9E8D57u, c.z, hi1);
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}

__device__ __forceinline__ float exp1_from_bits(uint32_t x) {
  const float u = ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1)
  return -__logf(u);
}

// one thread per environment: A is 6 .. 18 in every shipped configuration
__global__ void __launch_bounds__(128) sample_actions_kernel(const float* __restrict__ logits, int64_t n, int A,
                                                             const float* __restrict__ noise, uint64_t seed, uint64_t counter,
                                                             int64_t* __restrict__ actions, float* __restrict__ logp_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* l = logits + i * A;
  float m = -INFINITY;
  for (int a = 0; a < A; ++a) m = fmaxf(m, l[a]);
  float s = 0.f;
  for (int a = 0; a < A; ++a) s += expf(l[a] - m);
  const float inv = 1.0f / s;
  float best = -1.f;
  int arg = 0;
  uint4 bits = make_uint4(0, 0, 0, 0);
  for (int a = 0; a < A; ++a) {
    float q;
    if (noise) {
      q = noise[i * A + a];
    } else {
      if ((a & 3) == 0)
        bits = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)(i >> 32) ^ ((uint32_t)(a >> 2) << 16), (uint32_t)counter, (uint32_t)(counter >> 32)),
                             make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const uint32_t x = (a & 3) == 0 ? bits.x : (a & 3) == 1 ? bits.y : (a & 3) == 2 ? bits.z : bits.w;
      q = exp1_from_bits(x);
    }
    const float score = expf(l[a] - m) * inv / q;      // probs / q (torch.multinomial)
    if (score > best) { best = score; arg = a; }       // first maximum, like argmax
  }
  actions[i] = arg;
  if (logp_out) logp_out[i] = l[arg] - m - logf(s);
}

int sample_actions_launch(const float* logits, int64_t n, int A, const float* noise, uint64_t seed, uint64_t counter,
                          int64_t* actions, float* logp_out, cudaStream_t st) {
  if (n == 0) return DRL_OK;
  sample_actions_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(logits, n, A, noise, seed, counter, actions, logp_out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// ---- rollout ring ------------------------------------------------------------------------------------------------------
// frames of one time step of all envs: src [n_envs][28224 B] -> dst[(e * pitch + t)][28224 B]; one CTA per (env, slice)
__global__ void __launch_bounds__(256) record_frames_kernel(uint8_t* __restrict__ store, const uint8_t* __restrict__ src, int n_envs,
                                                            int64_t pitch_steps, int t, int slices) {
  constexpr int kVec = kObsBytes / 16;                 // 1764 x 16 B
  const int e = blockIdx.x / slices, sl = blockIdx.x % slices;
  if (e >= n_envs) return;
  const uint4* s = reinterpret_cast<const uint4*>(src + (size_t)e * kObsBytes);
  uint4* d = reinterpret_cast<uint4*>(store + ((size_t)e * pitch_steps + t) * kObsBytes);
  const int per = (kVec + slices - 1) / slices, lo = sl * per, hi = min(kVec, lo + per);
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) d[i] = s[i];
}

// FrameStackWrapper (drl/envs/wrappers/stateless/frame_stack.py:50-59, num_stack = 4) on the device: the stack of slot t_dst is
// the stack of slot t_src shifted by one channel with the newest frame appended (channel 3 = newest, what np.concatenate of
// the deque gives), or four copies of the newest frame where the environment was reset.  One 32-bit word per pixel:
// out = (prev >> 8) | (newest << 24).  The environments then hand over 7 KB per step instead of 28 KB.
__global__ void __launch_bounds__(256) stack_frames_kernel(uint8_t* __restrict__ store, const uint8_t* __restrict__ newest,
                                                           const uint8_t* __restrict__ reset, int n_envs, int64_t pitch_steps, int t_src,
                                                           int t_dst, int slices) {
  constexpr int kPix4 = 84 * 84 / 4;                   // 4 pixels per thread step: one uint32 of new pixels, one uint4 of stacks
  const int e = blockIdx.x / slices, sl = blockIdx.x % slices;
  if (e >= n_envs) return;
  const uint32_t* nw = reinterpret_cast<const uint32_t*>(newest + (size_t)e * 84 * 84);
  const uint4* src = reinterpret_cast<const uint4*>(store + ((size_t)e * pitch_steps + t_src) * kObsBytes);
  uint4* dst = reinterpret_cast<uint4*>(store + ((size_t)e * pitch_steps + t_dst) * kObsBytes);
  const int mode = reset != nullptr ? reset[e] : 0;  // 0 shift + append, 1 reset (four copies), 2 leave this environment alone
  if (mode == 2) return;
  const bool fresh = mode == 1;
  const int per = (kPix4 + slices - 1) / slices, lo = sl * per, hi = min(kPix4, lo + per);
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const uint32_t n4 = nw[i];
    uint4 o;
    if (fresh) {
      o.x = (n4 & 0xffu) * 0x01010101u;
      o.y = ((n4 >> 8) & 0xffu) * 0x01010101u;
      o.z = ((n4 >> 16) & 0xffu) * 0x01010101u;
      o.w = (n4 >> 24) * 0x01010101u;
    } else {
      const uint4 p = src[i];
      o.x = (p.x >> 8) | (n4 << 24);
      o.y = (p.y >> 8) | ((n4 >> 8) << 24);
      o.z = (p.z >> 8) | ((n4 >> 16) << 24);
      o.w = (p.w >> 8) | ((n4 >> 24) << 24);
    }
    dst[i] = o;
  }
}

struct RecordScalars {
  int64_t* act_store; float* done_store; float* rew_store[DRL_MAX_REWARD_STREAMS];
  const int64_t* a_t; const float* d_t; const float* r_t[DRL_MAX_REWARD_STREAMS];
  int n_streams;
};

__global__ void __launch_bounds__(256) record_scalars_kernel(RecordScalars p, int n_envs, int64_t pitch_steps, int t) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_envs) return;
  const int64_t row = (int64_t)e * pitch_steps + t;
  if (p.a_t) p.act_store[row] = p.a_t[e];
  if (p.d_t) p.done_store[row] = p.d_t[e];
  for (int k = 0; k < p.n_streams; ++k)
    if (p.r_t[k]) p.rew_store[k][row] = p.r_t[k][e];
}

// `count` consecutive time steps of every env from one [n_envs, pitch, step_bytes] array to another (step_bytes % 4 == 0)
__global__ void __launch_bounds__(256) copy_steps_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n_envs,
                                                         int64_t dst_pitch, int64_t src_pitch, int64_t words_per_step, int dst_t0, int src_t0,
                                                         int count) {
  const int64_t per_env = words_per_step * count;
  const int64_t total = per_env * n_envs;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i / per_env, w = i - e * per_env;
    dst[(e * dst_pitch + dst_t0) * words_per_step + w] = src[(e * src_pitch + src_t0) * words_per_step + w];
  }
}

}  // namespace drl

using namespace drl;

extern "C" int drl_sample_actions_f32(const float* logits, int64_t n, int num_actions, const float* noise, uint64_t seed,
                                      uint64_t counter, int64_t* actions_out, float* logp_out, void* stream) {
  if (!logits || !actions_out) return DRL_ERR_BAD_ARG;
  if (n < 0 || num_actions < 1 || num_actions > kMaxSampleActions) return DRL_ERR_BAD_SHAPE;
  return sample_actions_launch(logits, n, num_actions, noise, seed, counter, actions_out, logp_out, as_stream(stream));
}

extern "C" int drl_net_act(drl_net_t* net, const uint8_t* obs, const int64_t* row_idx, int nb, const float* noise, uint64_t seed,
                           uint64_t counter, int64_t* actions_out, float* logp_out, float* logits_out, float* values_out,
                           void* stream) {
  if (!net || !obs || !actions_out) return DRL_ERR_BAD_ARG;
  if (nb < 1 || nb > net->max_batch) return DRL_ERR_BAD_SHAPE;
  if (net->A > kMaxSampleActions) return DRL_ERR_BAD_SHAPE;
  // logits / values the caller does not want live in the (otherwise idle) d(pre-activation) workspace: >= 1024 B per row
  float* scratch = static_cast<float*>(net->dpre);
  float* logits = logits_out ? logits_out : scratch;
  float* values = values_out ? values_out : scratch + (size_t)nb * net->A;
  DRL_TRY(drl_net_forward(net, obs, row_idx, nb, logits, values, nullptr, nullptr, nullptr, stream));
  return sample_actions_launch(logits, nb, net->A, noise, seed, counter, actions_out, logp_out, as_stream(stream));
}

extern "C" int drl_rollout_record(uint8_t* obs_store, int64_t* act_store, float* done_store, float* const* reward_stores,
                                  int n_streams, int n_envs, int64_t pitch_steps, int t, const uint8_t* obs_t, const int64_t* a_t,
                                  const float* d_t, const float* const* r_t, int obs_on_host, void* stream) {
  if (n_envs < 1 || pitch_steps < 1 || t < 0 || t >= pitch_steps) return DRL_ERR_BAD_SHAPE;
  if (n_streams < 0 || n_streams > DRL_MAX_REWARD_STREAMS) return DRL_ERR_BAD_SHAPE;
  if ((obs_t && !obs_store) || (a_t && !act_store) || (d_t && !done_store)) return DRL_ERR_BAD_ARG;
  if (n_streams > 0 && r_t && !reward_stores) return DRL_ERR_BAD_ARG;
  cudaStream_t st = as_stream(stream);
  if (obs_t) {
    if ((reinterpret_cast<uintptr_t>(obs_t) | reinterpret_cast<uintptr_t>(obs_store)) & 15) return DRL_ERR_MISALIGNED;
    if (obs_on_host) {
      // straight into place: one strided host-to-device copy, no staging buffer on the device
      DRL_CUDA(cudaMemcpy2DAsync(obs_store + (size_t)t * kObsBytes, (size_t)pitch_steps * kObsBytes, obs_t, kObsBytes, kObsBytes,
                                 (size_t)n_envs, cudaMemcpyHostToDevice, st));
    } else {
      int slices = 1;                                  // >= 2 CTAs per SM in flight even for a handful of envs
      while (n_envs * slices < 2 * sm_count() && slices < 8) slices *= 2;
      record_frames_kernel<<<n_envs * slices, 256, 0, st>>>(obs_store, obs_t, n_envs, pitch_steps, t, slices);
      DRL_LAUNCH_CHECK();
    }
  }
  RecordScalars p{};
  p.act_store = act_store; p.done_store = done_store; p.a_t = a_t; p.d_t = d_t;
  bool any = a_t || d_t;
  if (r_t)
    for (int k = 0; k < n_streams; ++k) {
      p.rew_store[k] = reward_stores[k];
      p.r_t[k] = r_t[k];
      if (r_t[k] && !reward_stores[k]) return DRL_ERR_BAD_ARG;
      any = any || r_t[k];
    }
  p.n_streams = r_t ? n_streams : 0;
  if (any) {
    record_scalars_kernel<<<(n_envs + 255) / 256, 256, 0, st>>>(p, n_envs, pitch_steps, t);
    DRL_LAUNCH_CHECK();
  }
  return DRL_OK;
}

extern "C" int drl_rollout_copy_steps(void* dst, const void* src, int n_envs, int64_t dst_pitch_steps, int64_t src_pitch_steps,
                                      int64_t step_bytes, int dst_t0, int src_t0, int count, void* stream) {
  if (!dst || !src) return DRL_ERR_BAD_ARG;
  if (n_envs < 1 || count < 0 || dst_t0 < 0 || src_t0 < 0 || dst_t0 + count > dst_pitch_steps || src_t0 + count > src_pitch_steps)
    return DRL_ERR_BAD_SHAPE;
  if ((step_bytes & 3) || step_bytes < 4) return DRL_ERR_MISALIGNED;
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 3) return DRL_ERR_MISALIGNED;
  if (count == 0) return DRL_OK;
  const int64_t total = (step_bytes / 4) * count * n_envs;
  int blocks = (int)((total + 1023) / 1024);
  if (blocks > 4 * sm_count()) blocks = 4 * sm_count();
  copy_steps_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<uint32_t*>(dst), static_cast<const uint32_t*>(src), n_envs,
                                                           dst_pitch_steps, src_pitch_steps, step_bytes / 4, dst_t0, src_t0, count);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

extern "C" int drl_rollout_stack_frames(uint8_t* obs_store, int n_envs, int64_t pitch_steps, int t_src, int t_dst,
                                        const uint8_t* newest, const uint8_t* reset_mask, void* stream) {
  if (!obs_store || !newest) return DRL_ERR_BAD_ARG;
  if (n_envs < 1 || pitch_steps < 1 || t_src < 0 || t_dst < 0 || t_src >= pitch_steps || t_dst >= pitch_steps) return DRL_ERR_BAD_SHAPE;
  if ((reinterpret_cast<uintptr_t>(obs_store) & 15) || (reinterpret_cast<uintptr_t>(newest) & 3)) return DRL_ERR_MISALIGNED;
  int slices = 1;
  while (n_envs * slices < 2 * sm_count() && slices < 8) slices *= 2;
  stack_frames_kernel<<<n_envs * slices, 256, 0, as_stream(stream)>>>(obs_store, newest, reset_mask, n_envs, pitch_steps, t_src, t_dst, slices);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}
