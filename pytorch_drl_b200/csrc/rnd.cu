// RND predictor path (teacher / student networks, normalise+clip fused into the first conv).
// Placeholder entry points until the kernels land: they fail loudly, there is no fallback.
#include "common.cuh"
struct drl_rnd { int unused; };
extern "C" int drl_rnd_create(drl_rnd_t** out, int, int, int, int) { if (out) *out = nullptr; return DRL_ERR_UNSUPPORTED; }
extern "C" int drl_rnd_destroy(drl_rnd_t*) { return DRL_OK; }
extern "C" int64_t drl_rnd_param_count(int student, int widening) {
  const int64_t w = student ? widening : 1;
  int64_t n = 16 * w * 64 + 16 * w + 32 * w * 16 * w * 16 + 32 * w + 32 * w * 32 * w * 16 + 32 * w;
  if (student) n += 256 * w * 1152 * w + 256 * w + 256 * w * 256 * w + 256 * w + 512 * 256 * w + 512;
  else n += 512 * 1152 + 512;
  return n;
}
extern "C" int drl_rnd_set_params(drl_rnd_t*, const float*, const float*, void*) { return DRL_ERR_UNSUPPORTED; }
extern "C" int drl_rnd_reward(drl_rnd_t*, const uint8_t*, const int64_t*, int, const float*, const float*, float*, void*) { return DRL_ERR_UNSUPPORTED; }
extern "C" int drl_rnd_learn(drl_rnd_t*, const uint8_t*, const int64_t*, int, const float*, const float*, float*, float*, void*) { return DRL_ERR_UNSUPPORTED; }
