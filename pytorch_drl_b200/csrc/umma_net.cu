// tcgen05 bf16 engine of the NatureCNN trunk (DRL_PRECISION_BF16): weight packing into swizzled
// tile images, launch geometry of the forward / data-gradient / weight-gradient implicit GEMMs
// (umma_gemm.cuh), and the bias-gradient column sums.
//
// Layers (nature_cnn.py:27-36): conv1 8x8/4 (uint8 -> fp16 operands, 1/255 in the epilogue),
// conv2 4x4/2, conv3 3x3/1, fc 3136->512 (a 7x7 "conv" so that the NCHW flatten order
// c*49+h*7+w of nature_cnn.py:34 is honoured).  Activations are bf16 NHWC; features, logits,
// losses, gradients and Adam state are fp32.
#include "net.cuh"
#include <string.h>
#include <new>
#include "umma_gemm.cuh"
#include "umma_conv1.cuh"
#include "umma_tma_gemm.cuh"
#include "umma_tma_conv.cuh"
#include "umma_conv1_s2d.cuh"
#include "umma_tma_gemm2.cuh"
#include "umma_tma_conv2.cuh"

namespace drl {

static const float kInv255 = [] { const uint32_t u = 0x3b808081u; float f; memcpy(&f, &u, 4); return f; }();   // float32(1/255)

struct Bf16Engine {
  // packed B operands (tile images)
  void* w1f = nullptr;   // conv1 fwd   fp16 [4][32x64]
  void* w1s = nullptr;   // conv1 fwd   fp16 [4 taps][32x64] in space-to-depth block order (PK_CONV1_S2D)
  void* w2f = nullptr;   // conv2 fwd   bf16 [8][64x64]
  void* w3f = nullptr;   // conv3 fwd   bf16 [9][64x64]
  void* wff = nullptr;   // fc fwd      bf16 row-major [512][3136], k = (hw, c)  (TMA tensor map)
  void* wfd = nullptr;   // fc dgrad    bf16 row-major [3136][512] = the transpose      (TMA tensor map)
  void* w3d = nullptr;   // conv3 dgrad bf16 [9][64x64]   (flipped taps)
  uint32_t* bits[3] = {nullptr, nullptr, nullptr};   // ReLU masks of act1/act2/act3, 1 bit per element
  float* b1eff = nullptr; // conv1 effective bias (last block of pack_all_kernel)
  void* w2d = nullptr;   // conv2 dgrad bf16 [4][128x64]: rows = (parity class, ci)
  // weight-gradient scratch in GEMM order [kp = (ky, kx, ci)][co]: fc 3136 x 512, conv3 640 x 64, conv2 512 x 64.
  // All-zero between steps: wgrad_untranspose_kernel moves a layer's sums to the OIHW gradient and clears them.
  float* gt = nullptr;
};
constexpr int64_t kGtFc = 0, kGtConv3 = (int64_t)3136 * 512, kGtConv2 = kGtConv3 + 640 * 64, kGtTotal = kGtConv2 + 512 * 64;


// ---- packing -----------------------------------------------------------------------------------------
enum { PK_CONV_FWD = 0, PK_FC_FWD = 1, PK_FC_DGRAD = 2, PK_CONV3_DGRAD = 3, PK_CONV2_DGRAD = 4, PK_CONV1_S2D = 5 };

struct PackArgs {
  const float* w;      // reference-layout fp32 weights of the layer
  void* out;
  int mode, fp16;
  int variants, KB, N; // tile grid: [variants][KB][N rows x 64]
  int Cin, KH, KW, Cout;
};

__device__ __forceinline__ float pack_src(const PackArgs& a, int variant, int n, int k) {
  switch (a.mode) {
    case PK_CONV_FWD: {            // B[n=co][k=(ky,kx,ci)] = W[co,ci,ky,kx]
      const int kc = a.KW * a.Cin;
      const int ky = k / kc, r2 = k - ky * kc, kx = r2 / a.Cin, ci = r2 - kx * a.Cin;
      return a.w[((size_t)(n * a.Cin + ci) * a.KH + ky) * a.KW + kx];
    }
    case PK_FC_FWD: {              // B[n=variant*128+n][k=(hw,c)] = Wfc[n, c*49+hw]
      const int hw = k / 64, c = k - hw * 64;
      return a.w[(size_t)(variant * a.N + n) * 3136 + c * 49 + hw];
    }
    case PK_FC_DGRAD: {            // B[row = kfc = variant*64+n, (hw,c) order][k = n_fc] = Wfc[n_fc, c*49+hw]
      const int kfc = variant * 64 + n;
      const int hw = kfc / 64, c = kfc - hw * 64;
      return a.w[(size_t)k * 3136 + c * 49 + hw];
    }
    case PK_CONV3_DGRAD: {         // B[n=ci][k=(ty,tx,co)] = W3[co,ci,2-ty,2-tx]
      const int t = k / 64, co = k - t * 64, ty = t / 3, tx = t - ty * 3;
      return a.w[((size_t)(co * 64 + n) * 3 + (2 - ty)) * 3 + (2 - tx)];
    }
    case PK_CONV1_S2D: {           // B[n=co][k = tap*64 + e], tap = (ty,tx), e = dy*16 + px*4 + ci: W1[co,ci,4ty+dy,4tx+px]
      const int tap = k >> 6, e = k & 63, ty = tap >> 1, tx = tap & 1;
      const int dy = e >> 4, px = (e >> 2) & 3, ci = e & 3;
      return a.w[((size_t)(n * 4 + ci) * 8 + (4 * ty + dy)) * 8 + (4 * tx + px)];
    }
    default: {                     // PK_CONV2_DGRAD: row n = (class (py,px), ci); B[n][k=(ty,tx,co)] = W2[co,ci,py+2(1-ty),px+2(1-tx)]
      const int cls = n >> 5, ci = n & 31, py = cls >> 1, px = cls & 1;
      const int t = k / 64, co = k - t * 64, ty = t >> 1, tx = t & 1;
      return a.w[((size_t)(co * 32 + ci) * 4 + (py + 2 * (1 - ty))) * 4 + (px + 2 * (1 - tx))];
    }
  }
}

// ---- all operand layouts of one Adam step in ONE launch ---------------------------------------------------
// Block ranges: [0, fc_blocks) the two plain fc matrices, then one range per tile-image job, the last block
// computes conv1's effective bias.
struct PackAllArgs {
  PackArgs job[8];
  int first_block[9];     // job j owns blocks [first_block[j], first_block[j+1])
  int n_jobs;
  const float* wfc; __nv_bfloat16* fc_fwd; __nv_bfloat16* fc_dgrad; int fc_blocks;
  const float* w1; const float* b1; float scale; float* b1eff;
};

__global__ void __launch_bounds__(256) pack_all_kernel(const PackAllArgs a) {
  const int blk = blockIdx.x;
  if (blk < a.fc_blocks) {
    // fc weights [512][c*49+hw] -> forward [n][k = hw*64+c] and its transpose [k][n]; 32x32 tiles through shared
    // memory so both writes are coalesced
    __shared__ __nv_bfloat16 tile[32][33];
    const int tiles_k = 3136 / 32, n_tiles = 16 * tiles_k;
    for (int t = blk; t < n_tiles; t += a.fc_blocks) {
      const int tn = t / tiles_k, tk = t - tn * tiles_k;
      const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int n = tn * 32 + ty + 8 * r, k = tk * 32 + tx;
        const int hw = k >> 6, c = k & 63;
        const __nv_bfloat16 v = __float2bfloat16_rn(a.wfc[(size_t)n * 3136 + c * 49 + hw]);
        a.fc_fwd[(size_t)n * 3136 + k] = v;
        tile[ty + 8 * r][tx] = v;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int k = tk * 32 + ty + 8 * r, n = tn * 32 + tx;
        a.fc_dgrad[(size_t)k * 512 + n] = tile[tx][ty + 8 * r];
      }
      __syncthreads();
    }
    return;
  }
  const int jb = blk - a.fc_blocks;
  if (jb >= a.first_block[a.n_jobs]) {          // last block: conv1 effective bias, one warp per 4 output channels
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < 32; n += 8) {
      double s = 0.0;
      for (int k = lane; k < 256; k += 32) s += (double)__half2float(__float2half_rn(a.w1[n * 256 + k]));
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) a.b1eff[n] = (float)((double)a.b1[n] - 1024.0 * (double)a.scale * s);
    }
    return;
  }
  int j = 0;
  while (j + 1 < a.n_jobs && jb >= a.first_block[j + 1]) ++j;
  const PackArgs& pa = a.job[j];
  const int nblk = a.first_block[j + 1] - a.first_block[j], lb = jb - a.first_block[j];
  const int64_t total = (int64_t)pa.variants * pa.KB * pa.N * 64;
  for (int64_t e = (int64_t)lb * blockDim.x + threadIdx.x; e < total; e += (int64_t)nblk * blockDim.x) {
    const int kk = (int)(e & 63);
    const int n = (int)((e >> 6) % pa.N);
    const int64_t t = (e >> 6) / pa.N;
    const int kb = (int)(t % pa.KB), variant = (int)(t / pa.KB);
    const float v = pack_src(pa, variant, n, kb * 64 + kk);
    const int64_t tile_bytes = (int64_t)pa.N * 128;
    const int64_t off = t * tile_bytes + umma::swz_off<128>((uint32_t)n, (uint32_t)(kk >> 3)) + (kk & 7) * 2;
    if (pa.fp16) *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(pa.out) + off) = __float2half_rn(v);
    else *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(pa.out) + off) = __float2bfloat16_rn(v);
  }
}

// ---- bias gradients: column sums of a bf16 [R, C] matrix, C in {32, 64, 512} -----------------------------
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t R, int C,
                                                          float* __restrict__ out) {
  __shared__ float red[256 * 8];
  const int tpr = C / 8;                       // threads per row (each handles 8 columns = 16 B)
  const int rl = threadIdx.x / tpr, cg = threadIdx.x % tpr, rows_per_pass = 256 / tpr;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t r = (int64_t)blockIdx.x * rows_per_pass + rl; r < R; r += (int64_t)gridDim.x * rows_per_pass) {
    const uint4 v = umma::ldg_v4(x + r * C + cg * 8);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[2 * j] += __uint_as_float(w[j] << 16);
      acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < C) {
    const int cgi = threadIdx.x / 8, j = threadIdx.x % 8;
    float s = 0.f;
    for (int q = 0; q < rows_per_pass; ++q) s += red[(q * tpr + cgi) * 8 + j];
    atomicAdd(out + threadIdx.x, s);
  }
}

static int colsum(const void* x, int64_t R, int C, float* out, cudaStream_t st) {   // C in {32, 64}
  ProfileScope prof("bias_grad", st);
  const int tpr = C / 8;
  const int rows_per_pass = 256 / tpr;
  int blocks = (int)((R + rows_per_pass - 1) / rows_per_pass);
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  colsum_bf16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), R, C, out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// fc bias (512 columns): each thread owns 2 columns of a row pass
__global__ void __launch_bounds__(256) colsum512_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t R,
                                                             float* __restrict__ out) {
  float a0 = 0.f, a1 = 0.f;
  for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(x + r * 512 + threadIdx.x * 2);
    a0 += __uint_as_float(w << 16);
    a1 += __uint_as_float(w & 0xffff0000u);
  }
  atomicAdd(out + threadIdx.x * 2, a0);
  atomicAdd(out + threadIdx.x * 2 + 1, a1);
}

// ---- engine ---------------------------------------------------------------------------------------------
int bf16_alloc(drl_net* n) {
  Bf16Engine* e = new (std::nothrow) Bf16Engine();
  if (!e) return DRL_ERR_CUDA;
  n->bf16 = e;
  const size_t nb = (size_t)n->max_batch;
  DRL_CUDA(cudaMalloc(&n->act[0], nb * 12800 * 2));
  DRL_CUDA(cudaMalloc(&n->act[1], nb * 5184 * 2));
  DRL_CUDA(cudaMalloc(&n->act[2], nb * 3136 * 2));
  DRL_CUDA(cudaMalloc(&n->dact[0], nb * 12800 * 2));
  DRL_CUDA(cudaMalloc(&n->dact[1], nb * 5184 * 2));
  DRL_CUDA(cudaMalloc(&n->dact[2], nb * 3136 * 2));
  DRL_CUDA(cudaMalloc(&n->feat, nb * 512 * 4));
  DRL_CUDA(cudaMalloc(&n->dpre, nb * 512 * 2));
  DRL_CUDA(cudaMalloc(&e->w1f, 4 * 32 * 128));
  DRL_CUDA(cudaMalloc(&e->w1s, 4 * 32 * 128));
  DRL_CUDA(cudaMalloc(&e->w2f, 8 * 64 * 128));
  DRL_CUDA(cudaMalloc(&e->w3f, 9 * 64 * 128));
  DRL_CUDA(cudaMalloc(&e->wff, (size_t)512 * 3136 * 2));
  DRL_CUDA(cudaMalloc(&e->wfd, (size_t)3136 * 512 * 2));
  DRL_CUDA(cudaMalloc(&e->w3d, 9 * 64 * 128));
  DRL_CUDA(cudaMalloc(&e->w2d, 4 * 4 * 32 * 128));
  DRL_CUDA(cudaMalloc(&e->b1eff, 32 * sizeof(float)));
  DRL_CUDA(cudaMalloc(&e->gt, kGtTotal * sizeof(float)));
  DRL_CUDA(cudaMemset(e->gt, 0, kGtTotal * sizeof(float)));
  DRL_CUDA(cudaMalloc(&e->bits[0], nb * 12800 / 8));
  DRL_CUDA(cudaMalloc(&e->bits[1], nb * 5184 / 8));
  DRL_CUDA(cudaMalloc(&e->bits[2], nb * 3136 / 8));
  return DRL_OK;
}

void bf16_free(drl_net* n) {
  Bf16Engine* e = n->bf16;
  if (!e) return;
  cudaFree(e->w1f); cudaFree(e->w1s); cudaFree(e->w2f); cudaFree(e->w3f); cudaFree(e->wff); cudaFree(e->wfd); cudaFree(e->w3d); cudaFree(e->w2d); cudaFree(e->b1eff); cudaFree(e->gt); cudaFree(e->bits[0]); cudaFree(e->bits[1]); cudaFree(e->bits[2]);
  delete e;
  n->bf16 = nullptr;
}

int bf16_pack_params(drl_net* n, cudaStream_t st) {
  ProfileScope prof("pack_weights", st);
  Bf16Engine* e = n->bf16;
  const float* p = n->params;
  const int64_t* o = n->off;
  PackAllArgs a{};
  int nj = 0, blocks = 0;
  auto add = [&](const float* w, void* out, int mode, int fp16, int variants, int KB, int N, int Cin, int KH, int KW, int Cout) {
    a.job[nj] = PackArgs{w, out, mode, fp16, variants, KB, N, Cin, KH, KW, Cout};
    a.first_block[nj] = blocks;
    const int64_t total = (int64_t)variants * KB * N * 64;
    int b = (int)((total + 1023) / 1024);               // ~4 elements per thread
    blocks += b < 1 ? 1 : b;
    ++nj;
  };
  add(p + o[0], e->w1f, PK_CONV_FWD, 1, 1, 4, 32, 4, 8, 8, 32);
  add(p + o[0], e->w1s, PK_CONV1_S2D, 1, 1, 4, 32, 4, 8, 8, 32);
  add(p + o[2], e->w2f, PK_CONV_FWD, 0, 1, 8, 64, 32, 4, 4, 64);
  add(p + o[4], e->w3f, PK_CONV_FWD, 0, 1, 9, 64, 64, 3, 3, 64);
  add(p + o[4], e->w3d, PK_CONV3_DGRAD, 0, 1, 9, 64, 64, 3, 3, 64);
  add(p + o[2], e->w2d, PK_CONV2_DGRAD, 0, 1, 4, 128, 32, 4, 4, 64);
  a.first_block[nj] = blocks;
  a.n_jobs = nj;
  a.wfc = p + o[6]; a.fc_fwd = (__nv_bfloat16*)e->wff; a.fc_dgrad = (__nv_bfloat16*)e->wfd; a.fc_blocks = sm_count() * 2;
  a.w1 = p + o[0]; a.b1 = p + o[1]; a.scale = kInv255; a.b1eff = e->b1eff;
  pack_all_kernel<<<a.fc_blocks + blocks + 1, 256, 0, st>>>(a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

template <int N, int SRC, int EPI>
static int launch_rowgather(const char* site, const RowGatherArgs& a, int grid_y, cudaStream_t st) {
  ProfileScope prof(site, st);
  auto kern = rowgather_gemm_kernel<N, SRC, EPI>;
  constexpr int smem = RowGatherSmem<N>::kTotal;
  static bool configured = false;
  if (!configured) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles = (a.M + 127) / 128;
  int gx = 2 * sm_count() / grid_y;        // two CTAs per SM (96 KB smem, <= 256 TMEM columns each)
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  kern<<<dim3(gx, grid_y), 288, smem, st>>>(a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

template <int NT, int SRC, int MB>
static int launch_wgrad(const char* site, const WgradArgs& a, int grid_x, int grid_y, cudaStream_t st) {
  ProfileScope prof(site, st);
  auto kern = wgrad_gemm_kernel<NT, SRC, MB>;
  using Cf = WgradCfg<NT>;
  constexpr int smem = Cf::kStages * (2 * MB * Cf::kAtomBytes + Cf::kDyBytes) + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  kern<<<dim3(grid_x, grid_y), 288, smem, st>>>(a);
  DRL_LAUNCH_CHECK();
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

// [0] descriptor base-offset field for shifted operands: MEASURED on B200 (scripts/try_base_offset.py): the
//     128B swizzle is applied to the absolute smem address bits, so shifted starts need base offset 0.
//     [0] = 2: staging buffers of the TMA conv kernels packed at 128-byte instead of 1024-byte granularity (experiment
//     behind tma_conv2s_kernel: the TMA unit swizzles by absolute address bits too, scripts/try_unaligned_tma.py).
// [1] TMA implicit-GEMM conv kernels (1) or cp.async row-gather kernels (0)
// [2] conv weight-gradient kernels: 0 cp.async kernels, 1 TMA + shifted descriptors with one scattered atomic per
//     element, 2 (default) the same with 16-byte reductions into the [kp][co] scratch + wgrad_untranspose_kernel
//     (the fc weight gradient follows the same switch)
// [3] conv1 forward with its A operand in tensor memory (1) or in swizzled shared memory (0)
// [4] conv1 weight gradient on the space-to-depth block image (1) or the im2col kernel (0)
// [5] conv1 forward on the space-to-depth block image (1), else the kernel flag [3] selects
// [6] fc GEMMs: bits 0-1 CTA pairs (cta_group::2, 256 x 256 tiles): 0 none, 1 forward only (default), 2 forward + data
//     gradient; bit 2: the single-CTA data gradient with FOUR epilogue warpgroups (0.107 -> 0.102 ms) => default 5
// [7] descriptor-shifted conv kernels on CTA pairs (umma_tma_conv2.cuh), bit mask: 1 conv2 forward, 2 conv3 data
//     gradient, 4 conv2 data gradient.  Measured at 32 768 samples: conv2 fwd 0.255 -> 0.251 ms, conv3 dgrad 0.251 ->
//     0.232 ms, conv2 dgrad 0.253 -> 0.340 ms (it is HBM / epilogue bound: the pair only couples two epilogues) => 3.
//     bit 8: conv3 data gradient on CTA pairs with accumulator tiles that span samples (tma_conv2s_kernel, groups of 5
//     samples in 4 tiles): 0.235 -> 0.217 ms => default 11.
// [8] heads kernel compiled for 3 resident CTAs per SM (80 registers instead of 127, 64 bytes of spill) for the exact-(A, K)
//     training instantiations with A <= 8: 0.094 -> 0.084 ms => default 1
int g_debug_flags[16] = {0, 1, 2, 1, 1, 1, 5, 11, 1, 0, 0, 0, 0, 0, 0, 0};

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
  a.plane_bytes = g_debug_flags[0] == 2 ? a.box_bytes : (a.box_bytes + 1023) / 1024 * 1024;   // [0] = 2: experiment, 128-byte aligned TMA destinations
  a.raw_buf_bytes = a.plane_bytes * a.planes;
  const int window = (MROWS + max_shift) * 128;     // bytes a shifted 128-row window can reach from a plane start
  a.slack_bytes = window > a.plane_bytes ? (window - a.plane_bytes + 1023) / 1024 * 1024 : 0;
  if (a.MT != 1) return DRL_ERR_BAD_ARG;                // one 128-row tile per sample (all NatureCNN layers)
  if (a.bias_grad && a.bias_mod != (N == 64 ? 64 : 32)) return DRL_ERR_BAD_ARG;   // kernel's bias-group layout
  a.use_base_offset = 0;
  if (a.OH * a.RW > MROWS) return DRL_ERR_BAD_ARG;       // every output row must fall inside the tile
  if (a.KB != KB || a.KBX != KBX || a.SHIFT_RW != SHIFT_RW || a.planes != PLANES) return DRL_ERR_BAD_ARG;
  if constexpr (MROWS == 128) {
    if (pair_bit && (g_debug_flags[7] & pair_bit)) {     // CTA pairs: one sample per CTA, half of the weight rows each
      auto kern2 = tma_conv2_kernel<N, EPI, NR, NACC, NEPI, KB, KBX, SHIFT_RW, PLANES>;
      const int smem2 = tma_conv2_smem<N>(a.KB, a.raw_buf_bytes, NR, a.slack_bytes);
      static int configured2 = 0;
      if (configured2 < smem2) {
        DRL_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        configured2 = smem2;
      }
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
  static int configured = 0;
  if (configured < smem) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
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
  static bool configured = false;
  if (!configured) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
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
  static int configured = 0;
  if (configured < smem) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = smem;
  }
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
  if (BN == 256 && ((g_debug_flags[6] & 3) == 2 || ((g_debug_flags[6] & 3) == 1 && EPI == EPI_BIAS_RELU_F32))) {
    DRL_TRY(make_tmap_bf16(&tb, b, (uint64_t)N_total, (uint64_t)K, 128));
    if (args.bias_grad && args.bias_mod != 64) return DRL_ERR_BAD_ARG;
    auto kern2 = tma_gemm2_kernel<EPI>;
    static bool configured2 = false;
    if (!configured2) {
      DRL_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, kG2Smem));
      configured2 = true;
    }
    const int pair_tiles = ((M + 255) / 256) * ((N_total + 255) / 256);
    const int pairs = pair_tiles < sm_count() / 2 ? pair_tiles : sm_count() / 2;
    kern2<<<2 * pairs, kG2Threads, kG2Smem, st>>>(ta, tb, args);
    DRL_LAUNCH_CHECK();
    return DRL_OK;
  }
  DRL_TRY(make_tmap_bf16(&tb, b, (uint64_t)N_total, (uint64_t)K, BN));
  if (args.bias_grad && args.bias_mod != 64) return DRL_ERR_BAD_ARG;   // kernel's bias-group layout
  auto kern = tma_gemm_kernel<BN, EPI, NEPI>;
  constexpr int smem = TmaGemmCfg<BN>::kSmem;
  static bool configured = false;
  if (!configured) {
    DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int tiles = ((M + 127) / 128) * ((N_total + BN - 1) / BN);
  kern<<<tiles < sm_count() ? tiles : sm_count(), 32 * (4 * NEPI + 2), smem, st>>>(ta, tb, args);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

static RowGatherArgs fwd_args(int nb, const void* src, const int64_t* row_idx, int P, int OW, int H, int W, int C,
                              int stride, int KB, int KBX, const void* bpack, void* out, const float* bias,
                              float scale, int Ntotal) {
  RowGatherArgs a{};
  a.M = nb * P; a.KB = KB; a.src = src; a.row_idx = row_idx; a.P = P; a.OW = OW; a.H = H; a.W = W; a.C = C;
  a.stride = stride; a.pad = 0; a.KBX = KBX; a.bpack = bpack; a.b_variant_stride = 0; a.out = out; a.bias = bias;
  a.scale = scale; a.mask_src = nullptr; a.out_sample_pitch = (int64_t)P * Ntotal; a.out_row_pitch = OW * Ntotal;
  a.out_C = Ntotal; a.osy = 1; a.osx = 1; a.n_variant_stride = 0; a.classes = 0;
  return a;
}

int bf16_trunk_forward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, cudaStream_t st) {
  Bf16Engine* e = n->bf16;
  const float* p = n->params;
  const int64_t* o = n->off;
  // conv1: uint8 frames -> fp16 operands; y = relu(acc * (1/255) + b)   (scale_observations.py:35 fused)
  {
    ProfileScope prof("conv1_fwd", st);
    Conv1Args a1{};
    a1.obs = obs; a1.row_idx = row_idx; a1.nb = nb; a1.bpack = e->w1f; a1.bias = e->b1eff; a1.scale = kInv255;
    a1.out = reinterpret_cast<__nv_bfloat16*>(n->act[0]); a1.mask_bits = e->bits[0];
    static bool configured = false;
    if (!configured) {
      DRL_CUDA(cudaFuncSetAttribute(conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1FwdSmem));
      DRL_CUDA(cudaFuncSetAttribute(conv1_fwd_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1FwdTsSmem));
      DRL_CUDA(cudaFuncSetAttribute(conv1_fwd_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS2dFwdSmem));
      configured = true;
    }
    const int tiles = (nb * 400 + 127) / 128;
    if (g_debug_flags[5]) { // space-to-depth block image, every pixel converted once
      a1.bpack = e->w1s;
      conv1_fwd_s2d_kernel<<<nb < sm_count() ? nb : sm_count(), kS2dFThreads, kS2dFwdSmem, st>>>(a1);
    } else if (g_debug_flags[3])   // A operand staged in tensor memory (tcgen05.st + TS-form MMA)
      conv1_fwd_ts_kernel<<<tiles < sm_count() ? tiles : sm_count(), kC1tThreads, kC1FwdTsSmem, st>>>(a1);
    else
      conv1_fwd_kernel<<<tiles < sm_count() ? tiles : sm_count(), kC1Threads, kC1FwdSmem, st>>>(a1);
    DRL_LAUNCH_CHECK();
  }
  RowGatherArgs a2 = fwd_args(nb, n->act[0], nullptr, 81, 9, 20, 20, 32, 2, 8, 2, e->w2f, n->act[1], p + o[3], 1.f, 64);
  if (g_debug_flags[1]) {
    // act1 seen as [nb, 20 rows, 10 pixel pairs, 64]; even / odd rows are staged as two 10x10 planes (TMA row
    // stride 2), so tap row ky reads plane ky&1 shifted by (ky>>1)*10 + half and the GEMM rows are oy*10 + ox
    TmaConvArgs c2 = tma_conv_args(8, 2, 10, 10, 0, 1, 9, 9, e->w2f, n->act[1], p + o[3], nullptr, 5184, 9 * 64, 64, 1, 0);
    c2.planes = 2; c2.bits_out = e->bits[1];
    DRL_TRY((launch_tma_conv<64, EPI_BIAS_RELU_BF16, 4, 4, 2, 128, 8, 2, 10, 2>("conv2_fwd", n->act[0], nb, 20, 10, 10, 20, c2, st, 1)));
  } else {
    DRL_TRY((launch_rowgather<64, SRC_BF16, EPI_BIAS_RELU_BF16>("conv2_fwd", a2, 1, st)));
  }
  RowGatherArgs a3 = fwd_args(nb, n->act[1], nullptr, 49, 7, 9, 9, 64, 1, 9, 3, e->w3f, n->act[2], p + o[5], 1.f, 64);
  if (g_debug_flags[1]) {
    TmaConvArgs c3 = tma_conv_args(9, 3, 9, 9, 0, 1, 7, 7, e->w3f, n->act[2], p + o[5], nullptr, 3136, 7 * 64, 64, 1, 0);
    c3.bits_out = e->bits[2];
    DRL_TRY((launch_tma_conv<64, EPI_BIAS_RELU_BF16, 6, 4, 2, 64, 9, 3, 9, 1>("conv3_fwd", n->act[1], nb, 9, 9, 9, 9, c3, st)));
  } else {
    DRL_TRY((launch_rowgather<64, SRC_BF16, EPI_BIAS_RELU_BF16>("conv3_fwd", a3, 1, st)));
  }
  {
    TmaGemmArgs g{};
    g.M = nb; g.N_total = 512; g.KB = 49; g.out = n->feat; g.out_ld = 512; g.bias = p + o[7]; g.mask_src = nullptr;
    g.bias_grad = nullptr; g.bias_mod = 1; g.bits_in = nullptr;
    DRL_TRY((launch_tma_gemm<256, EPI_BIAS_RELU_F32>("fc_fwd", n->act[2], nb, 3136, e->wff, 512, g, st)));
  }
  return DRL_OK;
}

static WgradArgs wgrad_args(int nb, const void* src, const int64_t* row_idx, int P, int OW, int H, int W, int C,
                            int stride, int KBX, int atoms, const void* dy, int dy_ld, int mb_per_cta, int n_tiles,
                            float* gw, int Cin, int KH, int KW) {
  WgradArgs a{};
  a.M = nb * P; a.src = src; a.row_idx = row_idx; a.P = P; a.OW = OW; a.H = H; a.W = W; a.C = C; a.stride = stride;
  a.KBX = KBX; a.n_atoms_valid = atoms; a.dy = reinterpret_cast<const __nv_bfloat16*>(dy); a.dy_ld = dy_ld;
  a.mb_per_cta = mb_per_cta; a.n_tiles = n_tiles; a.gw = gw; a.Cin = Cin; a.KH = KH; a.KW = KW; a.scale = 1.f;
  return a;
}

int bf16_trunk_backward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, float* g, cudaStream_t st) {
  Bf16Engine* e = n->bf16;
  const int64_t* o = n->off;
  const int sms = sm_count();
  auto splits = [&](int M, int gy) { int s = (M + 31) / 32; int gx = (sms + gy - 1) / gy; return s < gx ? s : gx; };

  // ---- fc: weight gradient (3136 x 512, split over the batch), data gradient (masked by act3 > 0); the fc bias
  //      gradient was accumulated by the heads kernel (HeadParams::gfc_bias)
  {
    ProfileScope prof("fc_wgrad", st);
    using Cf = WgradTmaCfg<256, 1>;
    CUtensorMap tp, td;
    DRL_TRY(make_tmap_bf16(&tp, n->act[2], (uint64_t)nb, 3136, 32));
    DRL_TRY(make_tmap_bf16(&td, n->dpre, (uint64_t)nb, 512, 32));
    WgradTmaArgs w{};
    w.M = nb; w.n_tiles = 2; w.kp_valid = 3136; w.gw = g + o[6]; w.Cin = 64; w.KH = 7; w.KW = 7;
    w.gt = g_debug_flags[2] == 2 ? e->gt + kGtFc : nullptr; w.gt_ld = 512;
    auto kern = wgrad_tma_kernel<256, 1>;
    static bool configured = false;
    if (!configured) {
      DRL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::kSmem));
      configured = true;
    }
    int gx = 2 * sms / 50;                       // two CTAs per SM
    if (gx > (nb + 31) / 32) gx = (nb + 31) / 32;
    if (gx < 1) gx = 1;
    kern<<<dim3(gx, 50), 192, Cf::kSmem, st>>>(tp, td, w);
    DRL_LAUNCH_CHECK();
  }
  if (g_debug_flags[2] == 2) DRL_TRY(wgrad_untranspose("fc_wgrad_t", e->gt + kGtFc, g + o[6], 64, 49, 512, st));
  // fc weights, fc bias and the heads (everything from the fc weight on = 95 % of the gradient) are final: start
  // their all-reduce on the side stream while the convolution gradients are computed
  if (n->comm && comm_world(n->comm) > 1) DRL_TRY(comm_allreduce_after(n->comm, g, o[6], n->nparams, st));
  {
    TmaGemmArgs a{};
    a.M = nb; a.N_total = 3136; a.KB = 8; a.out = n->dact[2]; a.out_ld = 3136; a.bias = nullptr;
    a.mask_src = (const __nv_bfloat16*)n->act[2];
    a.bias_grad = g + o[5]; a.bias_mod = 64;          // conv3 bias gradient = column sums of dact3 (column % 64)
    a.bits_in = g_debug_flags[1] ? e->bits[2] : nullptr;   // bits are written by the TMA conv3 forward
    if (g_debug_flags[6] & 4) DRL_TRY((launch_tma_gemm<256, EPI_MASK_BF16, 4>("fc_dgrad", n->dpre, nb, 512, e->wfd, 3136, a, st)));   // 4 epilogue warpgroups
    else DRL_TRY((launch_tma_gemm<256, EPI_MASK_BF16>("fc_dgrad", n->dpre, nb, 512, e->wfd, 3136, a, st)));
  }
  // ---- conv3 (its bias gradient was accumulated by the fc data-gradient epilogue)
  WgradArgs w3 = wgrad_args(nb, n->act[1], nullptr, 49, 7, 9, 9, 64, 1, 3, 9, n->dact[2], 64, 5, 1, g + o[4], 64, 3, 3);
  if (g_debug_flags[2]) {
    // taps t = (ty, tx) shift the staged 9x9 input by ty*9 + tx rows; blocks = tap pairs (0,1)(2,3)(4,5)(6,7)(8,-)
    WgradConvArgs w{};
    w.ksteps = 4; w.n_blocks = 5; w.p_planes = 1; w.p_pad = 0;
    const int sh[10] = {0, 1, 2, 9, 10, 11, 18, 19, 20, 21};
    for (int mb = 0; mb < 5; ++mb) { w.a_off[mb] = sh[2 * mb] * 128; w.a_lbo[mb] = (sh[2 * mb + 1] - sh[2 * mb]) * 128; }
    w.p_plane_bytes = 11264;                         // 88 rows >= 63 + 21 + 1
    w.gw = g + o[4]; w.Cin = 64; w.KH = 3; w.KW = 3; w.kp_valid = 576;
    w.gt = g_debug_flags[2] == 2 ? e->gt + kGtConv3 : nullptr;
    DRL_TRY((launch_wgrad_conv<64, 5, 4, 4>("conv3_wgrad", n->act[1], 9, 9, 9, 9, n->dact[2], 7, 7, 9, 8, nb, w, st)));
    if (w.gt) DRL_TRY(wgrad_untranspose("conv3_wgrad_t", w.gt, g + o[4], 64, 9, 64, st));
  } else {
    DRL_TRY((launch_wgrad<64, SRC_BF16, 5>("conv3_wgrad", w3, splits(nb * 49, 1), 1, st)));
  }
  {
    RowGatherArgs a{};
    a.M = nb * 81; a.KB = 9; a.src = n->dact[2]; a.row_idx = nullptr; a.P = 81; a.OW = 9; a.H = 7; a.W = 7; a.C = 64;
    a.stride = 1; a.pad = 2; a.KBX = 3; a.bpack = e->w3d; a.b_variant_stride = 0;
    a.out = n->dact[1]; a.mask_src = (const __nv_bfloat16*)n->act[1]; a.scale = 1.f; a.bias = nullptr;
    a.out_sample_pitch = 5184; a.out_row_pitch = 9 * 64; a.out_C = 64; a.osy = 1; a.osx = 1;
    a.n_variant_stride = 0; a.classes = 0;
    if (g_debug_flags[1]) {
      // dY3 zero-padded by 2 (TMA out-of-bounds fill) to 11x11; output rows y*11 + x over the 9x9 input grid
      TmaConvArgs c = tma_conv_args(9, 3, 11, 11, 2, 1, 9, 9, e->w3d, n->dact[1], nullptr, n->act[1], 5184, 9 * 64, 64, 1, 0);
      c.bias_grad = g + o[3]; c.bias_mod = 64;       // conv2 bias gradient
      c.bits_in = e->bits[1];
      if (g_debug_flags[7] & 8)   // groups of 5 samples on one row stream: 4 accumulator tiles instead of 5
        DRL_TRY((launch_tma_conv2s<64, EPI_MASK_BF16, 2, 2, 9, 3, 11, 5, 9, 4>("conv3_dgrad", n->dact[2], nb, 7, 7, 11, c, st)));
      else
        DRL_TRY((launch_tma_conv<64, EPI_MASK_BF16, 4, 4, 2, 128, 9, 3, 11, 1>("conv3_dgrad", n->dact[2], nb, 7, 7, 11, 11, c, st, 2)));
    } else {
      DRL_TRY((launch_rowgather<64, SRC_BF16, EPI_MASK_BF16>("conv3_dgrad", a, 1, st)));
    }
  }
  // ---- conv2
  if (!g_debug_flags[1]) DRL_TRY(colsum(n->dact[1], (int64_t)nb * 81, 64, g + o[3], st));
  WgradArgs w2 = wgrad_args(nb, n->act[0], nullptr, 81, 9, 20, 20, 32, 2, 2, 8, n->dact[1], 64, 4, 1, g + o[2], 32, 4, 4);
  if (g_debug_flags[2]) {
    // act1 as pixel pairs, even / odd rows in two planes; block ky = atoms (ky, half 0/1): plane ky&1, shift (ky>>1)*10 + half
    WgradConvArgs w{};
    w.ksteps = 6; w.n_blocks = 4; w.p_planes = 2; w.p_pad = 0;
    w.p_plane_bytes = 14336;                         // 112 rows >= 95 + 11 + 1
    for (int ky = 0; ky < 4; ++ky) { w.a_off[ky] = (ky & 1) * w.p_plane_bytes + (ky >> 1) * 10 * 128; w.a_lbo[ky] = 128; }
    w.gw = g + o[2]; w.Cin = 32; w.KH = 4; w.KW = 4; w.kp_valid = 512;
    w.gt = g_debug_flags[2] == 2 ? e->gt + kGtConv2 : nullptr;
    DRL_TRY((launch_wgrad_conv<64, 4, 4, 6>("conv2_wgrad", n->act[0], 20, 10, 10, 20, n->dact[1], 9, 9, 10, 10, nb, w, st)));
    if (w.gt) DRL_TRY(wgrad_untranspose("conv2_wgrad_t", w.gt, g + o[2], 32, 16, 64, st));
  } else {
    DRL_TRY((launch_wgrad<64, SRC_BF16, 4>("conv2_wgrad", w2, splits(nb * 81, 1), 1, st)));
  }
  {
    // the four parity classes of the stride-2 transposed convolution share their 2x2-tap gather over dY2:
    // ONE GEMM with N = 4 classes x 32 ci; column chunk `cls` lands on output pixel (2u + cls/2, 2v + cls%2)
    RowGatherArgs a{};
    a.M = nb * 100; a.KB = 4; a.src = n->dact[1]; a.row_idx = nullptr; a.P = 100; a.OW = 10; a.H = 9; a.W = 9; a.C = 64;
    a.stride = 1; a.pad = 1; a.KBX = 2; a.bpack = e->w2d; a.b_variant_stride = 0;
    a.out = n->dact[0]; a.mask_src = (const __nv_bfloat16*)n->act[0]; a.scale = 1.f; a.bias = nullptr;
    a.out_sample_pitch = 12800; a.out_row_pitch = 20 * 32; a.out_C = 32; a.osy = 2; a.osx = 2;
    a.n_variant_stride = 0; a.classes = 1;
    if (g_debug_flags[1]) {
      // dY2 zero-padded by 1 to 11x11; GEMM rows u*11 + v over the 10x10 grid of 2x2 output blocks
      TmaConvArgs c = tma_conv_args(4, 2, 11, 11, 1, 1, 10, 10, e->w2d, n->dact[0], nullptr, n->act[0], 12800, 20 * 32, 32, 2, 1);
      c.bias_grad = g + o[1]; c.bias_mod = 32;       // conv1 bias gradient
      c.bits_in = e->bits[0];
      // FOUR epilogue warpgroups: with N = 128 (4 column chunks per row: mask, bf16 pack, 4 x 32 column sums) this layer
      // was bound by epilogue throughput — 2 / 3 / 4 warpgroups: 0.257 / 0.231 / 0.220 ms (the N = 64 layers do not gain)
      DRL_TRY((launch_tma_conv<128, EPI_MASK_BF16, 4, 4, 4, 128, 4, 2, 11, 1>("conv2_dgrad", n->dact[1], nb, 9, 9, 11, 11, c, st, 4)));
    } else {
      DRL_TRY((launch_rowgather<128, SRC_BF16, EPI_MASK_BF16>("conv2_dgrad", a, 1, st)));
    }
  }
  // ---- conv1 (no data gradient: observations need none, agent.py:66-67)
  if (!g_debug_flags[1]) DRL_TRY(colsum(n->dact[0], (int64_t)nb * 400, 32, g + o[1], st));
  if (g_debug_flags[4]) {
    ProfileScope prof("conv1_wgrad", st);
    CUtensorMap tdy;
    DRL_TRY(make_tmap_dy1(&tdy, n->dact[0], (uint64_t)nb));
    Conv1S2dArgs a{};
    a.obs = obs; a.row_idx = row_idx; a.nb = nb; a.scale = kInv255; a.gw = g + o[0];
    static bool configured = false;
    if (!configured) {
      DRL_CUDA(cudaFuncSetAttribute(conv1_wgrad_s2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kS2dWgradSmem));
      configured = true;
    }
    conv1_wgrad_s2d_kernel<<<nb < sms ? nb : sms, kS2dThreads, kS2dWgradSmem, st>>>(tdy, a);
    DRL_LAUNCH_CHECK();
  } else {
    ProfileScope prof("conv1_wgrad", st);
    Conv1Args a1{};
    a1.obs = obs; a1.row_idx = row_idx; a1.nb = nb; a1.scale = kInv255;
    a1.dy = reinterpret_cast<const __nv_bfloat16*>(n->dact[0]); a1.gw = g + o[0];
    static bool configured = false;
    if (!configured) {
      DRL_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kC1WgradSmem));
      configured = true;
    }
    const int tiles = (nb * 400 + 127) / 128;
    conv1_wgrad_kernel<<<tiles < sms ? tiles : sms, kC1Threads, kC1WgradSmem, st>>>(a1);
    DRL_LAUNCH_CHECK();
  }
  return DRL_OK;
}

}  // namespace drl

// debug switches (parity experiments): key 0 = descriptor base offset for shifted operands,
// key 1 = TMA implicit-GEMM conv kernels (1) or cp.async row-gather kernels (0)
extern "C" int drl_debug_set(int key, int value) {
  if (key < 0 || key >= 16) return DRL_ERR_BAD_ARG;
  drl::g_debug_flags[key] = value;
  return DRL_OK;
}
