#pragma once
#include <cuda_runtime.h>

namespace drl {
void count_launch();
// RAII: records a CUDA event pair around the launches of one site when profiling is enabled.
class ProfileScope {
 public:
  ProfileScope(const char* name, cudaStream_t st);
  ~ProfileScope();
 private:
  const char* name_;
  cudaStream_t st_;
  cudaEvent_t a_ = nullptr, b_ = nullptr;
  bool on_;
};
}  // namespace drl
