// Gradient exchange over NCCL (NVLink 5 / NVSwitch): replaces the DistributedDataParallel wrapper
// (drl/utils/launch.py:310; allreduce inside backward, drl/algos/ppo.py:233) and global_mean
// (drl/algos/common/metrics.py:8-26).  libnccl is resolved at run time with dlopen so that the
// library has no link-time dependency and shares the copy torch has already loaded.
#include <dlfcn.h>
#include <string.h>
#include <new>
#include <nccl.h>
#include "common.cuh"

namespace drl {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  bool ok = false;
};

static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) return nullptr;
#define SYM(field, name) *(void**)(&api.field) = dlsym(api.handle, name); if (!api.field) return nullptr
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
  api.ok = true;
  return &api;
}

}  // namespace drl

struct drl_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  cudaStream_t side = nullptr;          // overlapped gradient exchange (drl_net_set_comm)
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
};

using namespace drl;
static_assert(sizeof(ncclUniqueId) == DRL_NCCL_UNIQUE_ID_BYTES, "NCCL unique id size");

extern "C" int drl_comm_unique_id(uint8_t* id_host) {
  if (!id_host) return DRL_ERR_BAD_ARG;
  NcclApi* api = nccl_api();
  if (!api) return DRL_ERR_NCCL;
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return DRL_ERR_NCCL;
  memcpy(id_host, &id, sizeof(id));
  return DRL_OK;
}

extern "C" int drl_comm_create(drl_comm_t** out, const uint8_t* id_host, int rank, int world, int device) {
  if (!out || !id_host) return DRL_ERR_BAD_ARG;
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) return DRL_ERR_BAD_SHAPE;
  NcclApi* api = nccl_api();
  if (!api) return DRL_ERR_NCCL;
  DRL_TRY(drl_check_device(device));
  DRL_CUDA(cudaSetDevice(device));
  drl_comm* c = new (std::nothrow) drl_comm();
  if (!c) return DRL_ERR_NCCL;
  c->rank = rank; c->world = world; c->device = device;
  ncclUniqueId id;
  memcpy(&id, id_host, sizeof(id));
  if (api->CommInitRank(&c->comm, world, id, rank) != ncclSuccess) { delete c; return DRL_ERR_NCCL; }
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming) != cudaSuccess) {
    api->CommDestroy(c->comm);
    delete c;
    return DRL_ERR_CUDA;
  }
  *out = c;
  return DRL_OK;
}

extern "C" int drl_comm_destroy(drl_comm_t* c) {
  if (!c) return DRL_OK;
  NcclApi* api = nccl_api();
  if (api && c->comm) api->CommDestroy(c->comm);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ev_ready) cudaEventDestroy(c->ev_ready);
  if (c->ev_done) cudaEventDestroy(c->ev_done);
  delete c;
  return DRL_OK;
}

extern "C" int drl_grad_allreduce_bucketed(drl_comm_t* c, float* flat, const int64_t* bucket_offsets_host,
                                           int n_buckets, void* stream) {
  if (!c || !flat || !bucket_offsets_host) return DRL_ERR_BAD_ARG;
  if (n_buckets < 1) return DRL_ERR_BAD_SHAPE;
  NcclApi* api = nccl_api();
  if (!api) return DRL_ERR_NCCL;
  for (int b = 0; b < n_buckets; ++b) {
    const int64_t lo = bucket_offsets_host[b], hi = bucket_offsets_host[b + 1];
    if (hi < lo) return DRL_ERR_BAD_SHAPE;
    if (hi == lo) continue;
    if (api->AllReduce(flat + lo, flat + lo, (size_t)(hi - lo), ncclFloat, ncclSum, c->comm,
                       as_stream(stream)) != ncclSuccess)
      return DRL_ERR_NCCL;
  }
  return DRL_OK;
}

// ---- overlapped exchange used by the network handle (net.cu / umma_net.cu) ---------------------------------
namespace drl {

// SUM all-reduce of flat[lo, hi) on the communicator's side stream, ordered after everything `producer` has queued
int comm_allreduce_after(drl_comm* c, float* flat, int64_t lo, int64_t hi, cudaStream_t producer) {
  if (!c || hi <= lo) return DRL_OK;
  NcclApi* api = nccl_api();
  if (!api) return DRL_ERR_NCCL;
  DRL_CUDA(cudaEventRecord(c->ev_ready, producer));
  DRL_CUDA(cudaStreamWaitEvent(c->side, c->ev_ready, 0));
  if (api->AllReduce(flat + lo, flat + lo, (size_t)(hi - lo), ncclFloat, ncclSum, c->comm, c->side) != ncclSuccess)
    return DRL_ERR_NCCL;
  return DRL_OK;
}

// `consumer` waits for everything queued on the side stream so far
int comm_join(drl_comm* c, cudaStream_t consumer) {
  if (!c) return DRL_OK;
  DRL_CUDA(cudaEventRecord(c->ev_done, c->side));
  DRL_CUDA(cudaStreamWaitEvent(consumer, c->ev_done, 0));
  return DRL_OK;
}

int comm_world(const drl_comm* c) { return c ? c->world : 1; }

}  // namespace drl
