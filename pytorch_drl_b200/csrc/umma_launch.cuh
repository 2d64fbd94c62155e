// Host-side launch helpers shared by the tcgen05 engines (umma_net.cu: NatureCNN, umma_rnd.cu: RND predictor):
// TMA tensor maps (driver entry point resolved at run time), the launchers of the TMA conv / GEMM / weight-gradient
// kernel families, and the weight-gradient untranspose kernel.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "umma_gemm.cuh"
#include "umma_tma_gemm.cuh"
#include "umma_tma_conv.cuh"
#include "umma_tma_gemm2.cuh"
#include "umma_tma_conv2.cuh"

namespace drl {

extern int g_debug_flags[16];        // kernel-variant switches (umma_net.cu)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remember what each device was given.
constexpr int kMaxDevices = 64;
template <typename Kern>
static int ensure_dyn_smem(Kern kern, int smem, int (&configured)[kMaxDevices]) {
  int dev = 0;
  DRL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return DRL_ERR_BAD_ARG;
  if (configured[dev] < smem) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured[dev] = smem;
  }
  return DRL_OK;
}

// ---- TMA tensor maps (driver entry point resolved at run time: no link-time libcuda dependency) -----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major bf16 matrix [rows, cols] -> tiles of (64 columns = 128 B, box_rows rows), 128B swizzle, zero OOB fill
static int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return DRL_ERR_CUDA;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DRL_OK : DRL_ERR_CUDA;
}

// NHWC bf16 activation [nb, H, W, C64] (C64 = 64 elements per pixel row) -> 4-D tiles of (64, BW, BH, 1)
static int make_tmap_act4d(CUtensorMap* m, const void* base, uint64_t nb, uint64_t H, uint64_t W, uint32_t BW, uint32_t BH,
                           uint32_t h_stride = 1) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return DRL_ERR_CUDA;
  const cuuint64_t dims[4] = {64, W, H, nb};
  const cuuint64_t strides[3] = {128, W * 128, H * W * 128};
  const cuuint32_t box[4] = {64, BW, BH, 1};
  const cuuint32_t estr[4] = {1, 1, h_stride, 1};   // h_stride 2: every other map row (box BH spans BH/2 rows)
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DRL_OK : DRL_ERR_CUDA;
}

// dY1 [nb, 20, 20, 32] bf16 -> per sample a (32, 21, 20, 1) box: rows y*21 + x of 64 B, 64B swizzle; x = 20 is zero fill
static int make_tmap_dy1(CUtensorMap* m, const void* base, uint64_t nb) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return DRL_ERR_CUDA;
  const cuuint64_t dims[4] = {32, 20, 20, nb};
  const cuuint64_t strides[3] = {64, 20 * 64, 400 * 64};
  const cuuint32_t box[4] = {32, 21, 20, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DRL_OK : DRL_ERR_CUDA;
}

// uint8 frame store [rows, 84, 84, 4] seen as 32-bit words [rows][21 row groups][4 row phases][84 words]: a (84, 1, 21, 1) box is
// one ROW PHASE of a frame (image rows p, p + 4, ...) as 21 contiguous 336-byte rows, no swizzle (umma_conv1_i8.cuh).  The
// store's row count is not known here (rows are addressed through row_idx): the sample dimension is given its maximum.
static int make_tmap_frame_phases(CUtensorMap* m, const void* base) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return DRL_ERR_CUDA;
  const cuuint64_t dims[4] = {84, 4, 21, 1ull << 24};
  const cuuint64_t strides[3] = {336, 4 * 336, 84 * 336};
  const cuuint32_t box[4] = {84, 1, 21, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DRL_OK : DRL_ERR_CUDA;
}

template <int N, int EPI, int NR, int NACC, int NEPI, int MROWS, int KB, int KBX, int SHIFT_RW, int PLANES>
static int launch_tma_conv(const char* site, const void* src, int nb, int H, int W, int BW, int BH, TmaConvArgs a, cudaStream_t st,
                           int pair_bit = 0) {
  ProfileScope prof(site, st);
  CUtensorMap tm;
  DRL_TRY(make_tmap_act4d(&tm, src, (uint64_t)nb, (uint64_t)H, (uint64_t)W, (uint32_t)BW, (uint32_t)BH, (uint32_t)a.planes));
  a.nb = nb;
  a.box_bytes = BW * (BH / a.planes) * 128;
  const int tap_rows = (a.KB - 1) / a.KBX;               // last tap row
  const int max_shift = (a.planes == 2 ? tap_rows / 2 : tap_rows) * a.SHIFT_RW + (a.KBX - 1);
  a.plane_bytes = (a.box_bytes + 1023) / 1024 * 1024;
  a.raw_buf_bytes = a.plane_bytes * a.planes;
  const int window = (MROWS + max_shift) * 128;     // bytes a shifted 128-row window can reach from a plane start
  a.slack_bytes = window > a.plane_bytes ? (window - a.plane_bytes + 1023) / 1024 * 1024 : 0;
  if (a.MT != 1) return DRL_ERR_BAD_ARG;                // one 128-row tile per sample (all NatureCNN layers)
  if (a.bias_grad && (a.bias_mod < 1 || (N == 64 ? 64 : 32) % a.bias_mod != 0)) return DRL_ERR_BAD_ARG;   // kernel's bias-group layout
  a.use_base_offset = 0;
  if (a.OH * a.RW > MROWS) return DRL_ERR_BAD_ARG;       // every output row must fall inside the tile
  if (a.KB != KB || a.KBX != KBX || a.SHIFT_RW != SHIFT_RW || a.planes != PLANES) return DRL_ERR_BAD_ARG;
  if constexpr (MROWS == 128) {
    if (pair_bit && (g_debug_flags[7] & pair_bit)) {     // CTA pairs: one sample per CTA, half of the weight rows each
      auto kern2 = tma_conv2_kernel<N, EPI, NR, NACC, NEPI, KB, KBX, SHIFT_RW, PLANES>;
      const int smem2 = tma_conv2_smem<N>(a.KB, a.raw_buf_bytes, NR, a.slack_bytes);
      static int configured2[kMaxDevices] = {0};
      DRL_TRY(ensure_dyn_smem(kern2, smem2, configured2));
      const int total_pairs = (nb + 1) / 2;
      const int pairs = total_pairs < sm_count() / 2 ? total_pairs : sm_count() / 2;
      kern2<<<2 * pairs, 32 * (4 * NEPI + 2), smem2, st>>>(tm, a);
      DRL_LAUNCH_CHECK();
      return DRL_OK;
    }
  }
  auto kern = tma_conv_kernel<N, EPI, NR, NACC, NEPI, MROWS, KB, KBX, SHIFT_RW, PLANES>;
  const int smem = tma_conv_smem<N>(a.KB, a.raw_buf_bytes, NR, a.slack_bytes);
  if (smem > 227 * 1024) return DRL_ERR_BAD_ARG;
  static int configured[kMaxDevices] = {0};
  DRL_TRY(ensure_dyn_smem(kern, smem, configured));
  const int ctas = sm_count();
  kern<<<nb < ctas ? nb : ctas, 32 * (4 * NEPI + 2), smem, st>>>(tm, a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// CTA pairs + accumulator tiles that span samples (tma_conv2s_kernel): groups of G samples on one row stream
template <int N, int EPI, int NRG, int NEPI, int KB, int KBX, int SHIFT_RW, int G, int ROWS_PER_SAMPLE, int MT>
static int launch_tma_conv2s(const char* site, const void* src, int nb, int H, int W, int BW, TmaConvArgs a, cudaStream_t st) {
  ProfileScope prof(site, st);
  constexpr int SROWS = ROWS_PER_SAMPLE * SHIFT_RW;
  constexpr int MAX_SHIFT = ((KB - 1) / KBX) * SHIFT_RW + (KBX - 1);
  if (BW != SHIFT_RW || a.RW != SHIFT_RW || a.planes != 1 || a.KB != KB || a.KBX != KBX) return DRL_ERR_BAD_ARG;
  if (a.OH > ROWS_PER_SAMPLE || a.bias_grad && a.bias_mod != (N == 64 ? 64 : 32)) return DRL_ERR_BAD_ARG;
  CUtensorMap tm;
  DRL_TRY(make_tmap_act4d(&tm, src, (uint64_t)nb, (uint64_t)H, (uint64_t)W, (uint32_t)BW, (uint32_t)ROWS_PER_SAMPLE));
  a.nb = nb;
  a.box_bytes = SROWS * 128;
  auto kern = tma_conv2s_kernel<N, EPI, NRG, NEPI, KB, KBX, SHIFT_RW, G, SROWS, MT>;
  constexpr int smem = tma_conv2s_smem<N, NRG, KB, G, SROWS, MT, MAX_SHIFT>();
  static_assert(smem <= 227 * 1024, "shared memory plan too large");
  static int configured[kMaxDevices] = {0};
  DRL_TRY(ensure_dyn_smem(kern, smem, configured));
  const int total_pg = ((nb + G - 1) / G + 1) / 2;
  const int pairs = total_pg < sm_count() / 2 ? total_pg : sm_count() / 2;
  kern<<<2 * pairs, 32 * (4 * NEPI + 2), smem, st>>>(tm, a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

static TmaConvArgs tma_conv_args(int KB, int KBX, int RW, int SHIFT_RW, int pad, int MT, int OH, int OW, const void* bpack,
                                 void* out, const float* bias, const void* mask, int64_t sample_pitch, int row_pitch, int out_C,
                                 int os, int classes) {
  TmaConvArgs a{};
  a.KB = KB; a.KBX = KBX; a.RW = RW; a.SHIFT_RW = SHIFT_RW; a.pad = pad; a.planes = 1; a.MT = MT; a.OH = OH; a.OW = OW; a.bpack = bpack;
  a.out = out; a.bias = bias; a.scale = 1.f; a.mask_src = reinterpret_cast<const __nv_bfloat16*>(mask);
  a.out_sample_pitch = sample_pitch; a.out_row_pitch = row_pitch; a.out_C = out_C; a.osy = os; a.osx = os; a.classes = classes;
  return a;
}

// The weight-gradient kernels hold dW^T rows kp = (ky, kx, ci) with the output channels along the accumulator
// columns, the gradient buffer is OIHW (co slowest, (ci, ky, kx) contiguous): flushing the accumulators straight into
// it is one scattered 4-byte atomic per element from every CTA (measured: ~40 us of a 0.19 ms kernel).  Instead the
// CTAs reduce 16 bytes at a time into the [kp][co] scratch (full sectors), and this kernel moves the finished sums of one
// layer into the gradient buffer — both sides coalesced through a shared-memory tile — and clears the scratch.
//   grid (Cout / 32, Cin / CI_T), 256 threads; tile = rows {t * Cin + ci0 + cl} x 32 output channels.
template <int CI_T>
__global__ void __launch_bounds__(256) wgrad_untranspose_kernel(float* __restrict__ gt, float* __restrict__ g, int Cin, int khw,
                                                                int Cout) {
  extern __shared__ float tile[];                       // [32][CI_T * khw + 1]
  const int E = CI_T * khw, ES = E + 1;
  const int co0 = blockIdx.x * 32, ci0 = blockIdx.y * CI_T;
  constexpr int U = 8;                                  // loads in flight per thread (the kernel is latency bound)
  for (int base = threadIdx.x; base < E * 32; base += blockDim.x * U) {
    float v[U];
    float* src[U];
    int dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * (int)blockDim.x;
      const int row = idx >> 5, c = idx & 31;           // row = t * CI_T + cl
      const int t = row / CI_T, cl = row - t * CI_T;
      src[u] = gt + (int64_t)(t * Cin + ci0 + cl) * Cout + co0 + c;
      dst[u] = c * ES + cl * khw + t;
      v[u] = idx < E * 32 ? *src[u] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u * (int)blockDim.x < E * 32) {
        *src[u] = 0.f;
        tile[dst[u]] = v[u];
      }
    }
  }
  __syncthreads();
  for (int base = threadIdx.x; base < E * 32; base += blockDim.x * U) {
    float v[U];
    float* dstp[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * (int)blockDim.x;
      const int c = idx / E, e = idx - c * E;
      dstp[u] = g + ((int64_t)(co0 + c) * Cin + ci0) * khw + e;
      v[u] = idx < E * 32 ? *dstp[u] + tile[c * ES + e] : 0.f;     // += like the atomics it replaces
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (base + u * (int)blockDim.x < E * 32) *dstp[u] = v[u];
  }
}

// adds one layer's sums from the scratch to its OIHW gradient and leaves the scratch zero
static int wgrad_untranspose(const char* site, float* gt, float* g, int Cin, int khw, int Cout, cudaStream_t st) {
  ProfileScope prof(site, st);
  constexpr int CI_T = 2;
  const int smem = 32 * (CI_T * khw + 1) * (int)sizeof(float);
  wgrad_untranspose_kernel<CI_T><<<dim3(Cout / 32, Cin / CI_T), 256, smem, st>>>(gt, g, Cin, khw, Cout);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

template <int NT, int MB, int NR, int KSTEPS>
static int launch_wgrad_conv(const char* site, const void* p_src, int pH, int pW, int pBW, int pBH, const void* dy_src, int dH,
                             int dW, int dBW, int dBH, int nb, WgradConvArgs a, cudaStream_t st) {
  ProfileScope prof(site, st);
  CUtensorMap tp, td;
  DRL_TRY(make_tmap_act4d(&tp, p_src, (uint64_t)nb, (uint64_t)pH, (uint64_t)pW, (uint32_t)pBW, (uint32_t)pBH, (uint32_t)a.p_planes));
  DRL_TRY(make_tmap_act4d(&td, dy_src, (uint64_t)nb, (uint64_t)dH, (uint64_t)dW, (uint32_t)dBW, (uint32_t)dBH));
  a.nb = nb;
  a.p_box_bytes = pBW * (pBH / a.p_planes) * 128;
  a.dy_box_bytes = dBW * dBH * 128;
  a.p_buf_bytes = a.p_plane_bytes * a.p_planes;
  a.dy_buf_bytes = (a.dy_box_bytes + 1023) / 1024 * 1024;
  if (a.dy_buf_bytes < a.ksteps * 16 * 128) a.dy_buf_bytes = (a.ksteps * 16 * 128 + 1023) / 1024 * 1024;
  if (a.ksteps != KSTEPS) return DRL_ERR_BAD_ARG;
  auto kern = wgrad_conv_tma_kernel<NT, MB, NR, KSTEPS>;
  const int smem = NR * (a.p_buf_bytes + a.dy_buf_bytes) + 256 + 1024;
  static int configured[kMaxDevices] = {0};
  DRL_TRY(ensure_dyn_smem(kern, smem, configured));
  kern<<<nb < sm_count() ? nb : sm_count(), 192, smem, st>>>(tp, td, a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

template <int BN, int EPI, int NEPI = 2>
static int launch_tma_gemm(const char* site, const void* a, int M, int K, const void* b, int N_total, const TmaGemmArgs& args,
                           cudaStream_t st) {
  ProfileScope prof(site, st);
  CUtensorMap ta, tb;
  DRL_TRY(make_tmap_bf16(&ta, a, (uint64_t)M, (uint64_t)K, 128));
  // CTA pairs: each CTA stages 128 rows of A and 128 columns of B.  Measured: fc forward (49 k-blocks per tile)
  // 0.106 -> 0.098 ms, fc data gradient (8 k-blocks per tile, heavy epilogue) 0.105 -> 0.113 ms => [6] = 1 pairs the
  // forward only, [6] = 2 both.
  if (BN == 256 && ((g_debug_flags[6] & 3) == 2 || ((g_debug_flags[6] & 3) == 1 && (EPI == EPI_BIAS_RELU_F32 || EPI == EPI_BIAS_RELU_BF16)))) {
    DRL_TRY(make_tmap_bf16(&tb, b, (uint64_t)N_total, (uint64_t)K, 128));
    if (args.bias_grad && args.bias_mod != 64) return DRL_ERR_BAD_ARG;
    auto kern2 = tma_gemm2_kernel<EPI>;
    static int configured2[kMaxDevices] = {0};
    DRL_TRY(ensure_dyn_smem(kern2, kG2Smem, configured2));
    const int pair_tiles = ((M + 255) / 256) * ((N_total + 255) / 256);
    const int pairs = pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2;
    kern2<<<2 * pairs, kG2Threads, kG2Smem, st>>>(ta, tb, args);
    DRL_LAUNCH_CHECK();
    return DRL_OK;
  }
  DRL_TRY(make_tmap_bf16(&tb, b, (uint64_t)N_total, (uint64_t)K, BN));
  if (args.bias_grad && args.bias_mod != 64 && args.bias_mod != 32) return DRL_ERR_BAD_ARG;   // kernel's bias-group layout
  auto kern = tma_gemm_kernel<BN, EPI, NEPI>;
  constexpr int smem = TmaGemmCfg<BN>::kSmem;
  static int configured[kMaxDevices] = {0};
  DRL_TRY(ensure_dyn_smem(kern, smem, configured));
  const int tiles = ((M + 127) / 128) * ((N_total + BN - 1) / BN);
  kern<<<tiles < sm_count() ? tiles : sm_count(), 32 * (4 * NEPI + 2), smem, st>>>(ta, tb, args);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

}  // namespace drl
