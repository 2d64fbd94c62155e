// tcgen05 bf16 engine of the NatureCNN trunk — placeholder until the kernels land.
#include "net.cuh"
namespace drl {
int bf16_alloc(drl_net*) { return DRL_ERR_UNSUPPORTED; }
void bf16_free(drl_net*) {}
int bf16_pack_params(drl_net*, cudaStream_t) { return DRL_ERR_UNSUPPORTED; }
int bf16_trunk_forward(drl_net*, const uint8_t*, const int64_t*, int, cudaStream_t) { return DRL_ERR_UNSUPPORTED; }
int bf16_trunk_backward(drl_net*, const uint8_t*, const int64_t*, int, float*, cudaStream_t) { return DRL_ERR_UNSUPPORTED; }
}  // namespace drl
