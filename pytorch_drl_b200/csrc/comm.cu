// Gradient exchange over NCCL (NVLink 5 / NVSwitch): replaces the DistributedDataParallel wrapper
// (drl/utils/launch.py:310; allreduce inside backward, drl/algos/ppo.py:233) and global_mean
// (drl/algos/common/metrics.py:8-26).  libnccl is resolved at run time with dlopen so that the
// library has no link-time dependency and shares the copy torch has already loaded.
#include <dlfcn.h>
#include <string.h>
#include <new>
#include <nccl.h>
#include "common.cuh"
#include "profile.cuh"

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
  // ---- NVLink peer-memory all-reduce (drl_comm_p2p_export / _import) ----
  uint8_t* xbuf = nullptr;              // my exchange buffer (cudaMalloc, IPC-exported)
  int64_t x_elems = 0;                  // capacity of each of its 4 regions, in floats
  uint8_t* peer[DRL_MAX_P2P_RANKS] = {nullptr};   // every rank's exchange buffer mapped here ([rank] = xbuf)
  uint8_t** peer_dev = nullptr;         // the same table in device memory
  unsigned int* counters = nullptr;     // [2] last-CTA counters of the publish / reduce kernels
  uint32_t epoch = 0;
  bool p2p_ready = false;
};


// =============================================================================================================
// All-reduce over NVLink peer memory (two-shot: reduce-scatter + all-gather through IPC-mapped exchange
// buffers).  The flat gradient is 6.75 MB: at that size an NCCL ring/tree call is latency- and launch-bound
// (measured: ~0.2 ms per learner step on 2-8 B200s), while every GPU can simply READ its 1/N slice from all peers
// and WRITE the sums back at NVLink speed.  Exchange buffer of rank r (cudaMalloc + cudaIpc):
//     [ flags 256 B | data[2][cap] | out[2][cap] ]      (2 = parity of the call counter)
// flags: ready[q] at +0, done[q] at +64 (uint32, written by rank q with system-scope release stores).
// Call e (the same e on every rank; collective calls are issued in the same order everywhere):
//   publish  grads -> my data[e&1]; the last CTA stores ready[me] = e into EVERY rank's flags
//   reduce   wait ready[q] >= e for all q; my slice: sum over q = 0..N-1 of peer q's data[e&1] (fixed order =>
//            every rank ends with bit-identical sums), stored into every rank's out[e&1]; the last CTA
//            stores done[me] = e everywhere
//   collect  wait done[q] >= e for all q; my out[e&1] -> grads
// Parity double-buffering makes the protocol safe without a trailing barrier: a rank can run at most one call
// ahead of the slowest one (its next `reduce` waits for everybody's `publish` of that call).
// =============================================================================================================
namespace drl {

constexpr int kP2pFlagsBytes = 256;
constexpr int kP2pThreads = 256;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_v4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_volatile_f(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

struct P2pArgs {
  uint8_t** peer;        // [world] exchange buffers (device table)
  int rank, world;
  int64_t cap;           // floats per region
  int64_t n;             // floats of this call
  uint32_t epoch;
  unsigned int* counter; // last-CTA detection
  float* grads;
};

__device__ __forceinline__ float* p2p_data(uint8_t* base, int64_t cap, uint32_t parity) {
  return reinterpret_cast<float*>(base + kP2pFlagsBytes) + (int64_t)parity * cap;
}
__device__ __forceinline__ float* p2p_out(uint8_t* base, int64_t cap, uint32_t parity) {
  return reinterpret_cast<float*>(base + kP2pFlagsBytes) + (int64_t)(2 + parity) * cap;
}

// every CTA: fence; the last one to arrive publishes `epoch` in slot `slot` of every rank's flag block
__device__ __forceinline__ void p2p_signal_all(const P2pArgs& a, int slot_offset_bytes) {
  __shared__ bool last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    if (threadIdx.x < a.world)
      st_release_sys(reinterpret_cast<uint32_t*>(a.peer[threadIdx.x] + slot_offset_bytes) + a.rank, a.epoch);
    if (threadIdx.x == 0) *a.counter = 0;
  }
}
__device__ __forceinline__ void p2p_wait_all(const P2pArgs& a, int slot_offset_bytes) {
  if (threadIdx.x < a.world) {
    const uint32_t* f = reinterpret_cast<const uint32_t*>(a.peer[a.rank] + slot_offset_bytes) + threadIdx.x;
    while ((int32_t)(ld_acquire_sys(f) - a.epoch) < 0) __nanosleep(64);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kP2pThreads) p2p_publish_kernel(const P2pArgs a) {
  float* dst = p2p_data(a.peer[a.rank], a.cap, a.epoch & 1u);
  const int64_t n4 = a.n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(a.grads)[i];
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) dst[(n4 << 2) + threadIdx.x] = a.grads[(n4 << 2) + threadIdx.x];
  p2p_signal_all(a, 0);
}

__global__ void __launch_bounds__(kP2pThreads) p2p_reduce_kernel(const P2pArgs a) {
  p2p_wait_all(a, 0);
  const uint32_t par = a.epoch & 1u;
  // slices of whole float4s; the last rank also takes the scalar tail
  const int64_t n4 = a.n >> 2, per = (n4 + a.world - 1) / a.world;
  const int64_t lo = a.rank * per, hi = min(n4, lo + per);
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < a.world; ++q) {
      const float4 v = ld_volatile_v4(p2p_data(a.peer[q], a.cap, par) + (i << 2));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    for (int q = 0; q < a.world; ++q) reinterpret_cast<float4*>(p2p_out(a.peer[q], a.cap, par))[i] = s;
  }
  if (a.rank == a.world - 1 && blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float s = 0.f;
    for (int q = 0; q < a.world; ++q) s += ld_volatile_f(p2p_data(a.peer[q], a.cap, par) + i);
    for (int q = 0; q < a.world; ++q) p2p_out(a.peer[q], a.cap, par)[i] = s;
  }
  p2p_signal_all(a, 64);
}

__global__ void __launch_bounds__(kP2pThreads) p2p_collect_kernel(const P2pArgs a) {
  p2p_wait_all(a, 64);
  const float* src = p2p_out(a.peer[a.rank], a.cap, a.epoch & 1u);
  const int64_t n4 = a.n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(a.grads)[i] = ld_volatile_v4(src + (i << 2));
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) a.grads[(n4 << 2) + threadIdx.x] = ld_volatile_f(src + (n4 << 2) + threadIdx.x);
}

static int p2p_allreduce(drl_comm* c, float* grads, int64_t n, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(grads) & 15) != 0) return DRL_ERR_MISALIGNED;
  ProfileScope prof(n > 100000 ? "allreduce_fc_heads" : "allreduce_small", st);
  P2pArgs a{};
  a.peer = c->peer_dev; a.rank = c->rank; a.world = c->world; a.cap = c->x_elems; a.n = n;
  a.epoch = ++c->epoch; a.grads = grads;
  const int64_t n4 = n >> 2;
  int blocks = (int)((n4 + kP2pThreads * 4 - 1) / (kP2pThreads * 4));
  blocks = blocks < 1 ? 1 : (blocks > 96 ? 96 : blocks);
  int rblocks = (int)((n4 / c->world + kP2pThreads * 2 - 1) / (kP2pThreads * 2));
  rblocks = rblocks < 1 ? 1 : (rblocks > 96 ? 96 : rblocks);
  a.counter = c->counters;
  p2p_publish_kernel<<<blocks, kP2pThreads, 0, st>>>(a);
  DRL_LAUNCH_CHECK();
  a.counter = c->counters + 1;
  p2p_reduce_kernel<<<rblocks, kP2pThreads, 0, st>>>(a);
  DRL_LAUNCH_CHECK();
  p2p_collect_kernel<<<blocks, kP2pThreads, 0, st>>>(a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

}  // namespace drl

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
  for (int q = 0; q < c->world && q < DRL_MAX_P2P_RANKS; ++q)
    if (c->peer[q] && q != c->rank) cudaIpcCloseMemHandle(c->peer[q]);
  cudaFree(c->xbuf); cudaFree(c->peer_dev); cudaFree(c->counters);
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
    if (c->p2p_ready && hi - lo <= c->x_elems) {
      DRL_TRY(p2p_allreduce(c, flat + lo, hi - lo, as_stream(stream)));
      continue;
    }
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
  if (c->p2p_ready && hi - lo <= c->x_elems) return p2p_allreduce(c, flat + lo, hi - lo, c->side);
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

extern "C" int drl_comm_p2p_export(drl_comm_t* c, int64_t max_elems, uint8_t* handle_out) {
  if (!c || !handle_out || max_elems < 1) return DRL_ERR_BAD_ARG;
  if (c->world > DRL_MAX_P2P_RANKS) return DRL_ERR_UNSUPPORTED;
  if (c->xbuf) return DRL_ERR_BAD_ARG;                       // once per communicator
  DRL_CUDA(cudaSetDevice(c->device));
  c->x_elems = (max_elems + 63) / 64 * 64;                    // regions stay 256-byte aligned
  const size_t bytes = kP2pFlagsBytes + 4 * (size_t)c->x_elems * sizeof(float);
  DRL_CUDA(cudaMalloc(&c->xbuf, bytes));
  DRL_CUDA(cudaMemset(c->xbuf, 0, kP2pFlagsBytes));
  DRL_CUDA(cudaMalloc(&c->peer_dev, DRL_MAX_P2P_RANKS * sizeof(uint8_t*)));
  DRL_CUDA(cudaMalloc(&c->counters, 2 * sizeof(unsigned int)));
  DRL_CUDA(cudaMemset(c->counters, 0, 2 * sizeof(unsigned int)));
  DRL_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  DRL_CUDA(cudaIpcGetMemHandle(&h, c->xbuf));
  static_assert(sizeof(h) == DRL_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  return DRL_OK;
}

extern "C" int drl_comm_p2p_import(drl_comm_t* c, const uint8_t* handles_all) {
  if (!c) return DRL_ERR_BAD_ARG;
  if (!handles_all) { c->p2p_ready = false; return DRL_OK; }   // back to NCCL
  if (!c->xbuf) return DRL_ERR_BAD_ARG;
  DRL_CUDA(cudaSetDevice(c->device));
  for (int q = 0; q < c->world; ++q) {
    if (q == c->rank) { c->peer[q] = c->xbuf; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles_all + (size_t)q * DRL_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    DRL_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer[q] = static_cast<uint8_t*>(p);
  }
  DRL_CUDA(cudaMemcpy(c->peer_dev, c->peer, DRL_MAX_P2P_RANKS * sizeof(uint8_t*), cudaMemcpyHostToDevice));
  c->p2p_ready = true;
  return DRL_OK;
}
