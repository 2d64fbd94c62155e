// tcgen05 bf16 engine of the NatureCNN trunk (DRL_PRECISION_BF16): weight packing into swizzled
// tile images and the launch geometry of the forward / data-gradient / weight-gradient implicit GEMMs
// (launchers in umma_launch.cuh).
//
// Layers (nature_cnn.py:27-36): conv1 8x8/4 (uint8 -> fp16 operands, 1/255 in the epilogue),
// conv2 4x4/2, conv3 3x3/1, fc 3136->512 (a 7x7 "conv" so that the NCHW flatten order
// c*49+h*7+w of nature_cnn.py:34 is honoured).  Activations are bf16 NHWC; features, logits,
// losses, gradients and Adam state are fp32.
#include "net.cuh"
#include <string.h>
#include <new>
#include "umma_launch.cuh"
#include "umma_conv1_s2d.cuh"
#include "umma_heads.cuh"
#include "umma_conv1_i8.cuh"

namespace drl {

static const float kInv255 = [] { const uint32_t u = 0x3b808081u; float f; memcpy(&f, &u, 4); return f; }();   // float32(1/255)

struct Bf16Engine {
  // packed B operands (tile images)
  void* w2f = nullptr;   // conv2 fwd   bf16 [8][64x64]
  void* w3f = nullptr;   // conv3 fwd   bf16 [9][64x64]
  void* wff = nullptr;   // fc fwd      bf16 row-major [512][3136], k = (hw, c)  (TMA tensor map)
  void* wfd = nullptr;   // fc dgrad    bf16 row-major [3136][512] = the transpose      (TMA tensor map)
  void* w3d = nullptr;   // conv3 dgrad bf16 [9][64x64]   (flipped taps)
  uint32_t* bits[3] = {nullptr, nullptr, nullptr};   // ReLU masks of act1/act2/act3, 1 bit per element
  int8_t* w1q = nullptr;  // conv1 fwd, integer kernel: s8 [8 taps][64 rows = (hi, lo) x 32 co][32 B = (kx, ci)] core-matrix images (umma_conv1_i8.cuh)
  float* c1scale = nullptr; // [32] per-channel step of the 16-bit fixed-point weights x float32(1/255)
  void* w2d = nullptr;   // conv2 dgrad bf16 [4][128x64]: rows = (parity class, ci)
  void* whd = nullptr;   // heads       bf16 [8][32x64]: rows = (policy actions, value streams, zero padding)  (umma_heads.cuh)
  __nv_bfloat16* featb = nullptr;   // [nb, 512] bf16 features of the learner step (the fp32 `feat` serves the other entry points)
  // weight-gradient scratch in GEMM order [kp = (ky, kx, ci)][co]: fc 3136 x 512, conv3 640 x 64, conv2 512 x 64.
  // All-zero between steps: wgrad_untranspose_kernel moves a layer's sums to the OIHW gradient and clears them.
  float* gt = nullptr;
};
constexpr int64_t kGtFc = 0, kGtConv3 = (int64_t)3136 * 512, kGtConv2 = kGtConv3 + 640 * 64, kGtTotal = kGtConv2 + 512 * 64;


// ---- packing -----------------------------------------------------------------------------------------
enum { PK_CONV_FWD = 0, PK_FC_FWD = 1, PK_FC_DGRAD = 2, PK_CONV3_DGRAD = 3, PK_CONV2_DGRAD = 4, PK_HEADS = 6 };

struct PackArgs {
  const float* w;      // reference-layout fp32 weights of the layer
  void* out;
  int mode, fp16;
  int variants, KB, N; // tile grid: [variants][KB][N rows x 64]
  int Cin, KH, KW, Cout;
  const float* wv[DRL_MAX_REWARD_STREAMS];   // PK_HEADS: value-head weight rows (Cin = A, KH = K)
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
    case PK_HEADS:                 // B[n = head][k = feature]: policy rows, value rows, zero padding
      return n < a.Cin ? a.w[(size_t)n * 512 + k] : (n < a.Cin + a.KH ? a.wv[n - a.Cin][k] : 0.f);
    default: {                     // PK_CONV2_DGRAD: row n = (class (py,px), ci); B[n][k=(ty,tx,co)] = W2[co,ci,py+2(1-ty),px+2(1-tx)]
      const int cls = n >> 5, ci = n & 31, py = cls >> 1, px = cls & 1;
      const int t = k / 64, co = k - t * 64, ty = t >> 1, tx = t & 1;
      return a.w[((size_t)(co * 32 + ci) * 4 + (py + 2 * (1 - ty))) * 4 + (px + 2 * (1 - tx))];
    }
  }
}

// ---- all operand layouts of one Adam step in ONE launch ---------------------------------------------------
// Block ranges: [0, fc_blocks) the two plain fc matrices, then one range per tile-image job, the last block
// quantises conv1's weights for the integer forward kernel.
struct PackAllArgs {
  PackArgs job[8];
  int first_block[9];     // job j owns blocks [first_block[j], first_block[j+1])
  int n_jobs;
  const float* wfc; __nv_bfloat16* fc_fwd; __nv_bfloat16* fc_dgrad; int fc_blocks;
  const float* w1; float scale; int8_t* w1q; float* c1scale;      // conv1 forward: 16-bit fixed-point weights (umma_conv1_i8.cuh)
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
  if (jb >= a.first_block[a.n_jobs]) {          // last block: conv1's fixed-point weights, one warp per 4 output channels
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < 32; n += 8) {
      // integer conv1: channel n as 16-bit fixed point q * delta, q = 256 hi + lo (hi, lo in s8); B tile of tap ky:
      // byte (row, k = kx*4 + ci) at ky*2048 + (row/8)*256 + (k/16)*128 + (row%8)*16 + k%16, rows n (hi) and 32 + n (lo)
      float mx = 0.f;
      for (int k = lane; k < 256; k += 32) mx = fmaxf(mx, fabsf(a.w1[n * 256 + k]));
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float delta = mx > 0.f ? mx / 32639.f : 1.f;      // |q| <= 127 * 256 + 127: hi = (q + 128) >> 8 stays in s8
      for (int k = lane; k < 256; k += 32) {
        const int ci = k >> 6, ky = (k >> 3) & 7, kx = k & 7;
        int q = __float2int_rn(a.w1[n * 256 + k] / delta);
        q = q > 32639 ? 32639 : (q < -32639 ? -32639 : q);
        const int hi = (q + 128) >> 8, lo = q - hi * 256;
        const int kb = kx * 4 + ci;
        const int base = ky * 2048 + (kb >> 4) * 128 + (kb & 15);
        a.w1q[base + (n >> 3) * 256 + (n & 7) * 16] = (int8_t)hi;
        a.w1q[base + ((32 + n) >> 3) * 256 + ((32 + n) & 7) * 16] = (int8_t)lo;
      }
      if (lane == 0) a.c1scale[n] = (float)((double)delta * (double)a.scale);
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
  DRL_CUDA(cudaMalloc(&e->w2f, 8 * 64 * 128));
  DRL_CUDA(cudaMalloc(&e->w3f, 9 * 64 * 128));
  DRL_CUDA(cudaMalloc(&e->wff, (size_t)512 * 3136 * 2));
  DRL_CUDA(cudaMalloc(&e->wfd, (size_t)3136 * 512 * 2));
  DRL_CUDA(cudaMalloc(&e->w3d, 9 * 64 * 128));
  DRL_CUDA(cudaMalloc(&e->w2d, 4 * 4 * 32 * 128));
  DRL_CUDA(cudaMalloc(&e->whd, kHmWBytes));
  DRL_CUDA(cudaMalloc(&e->featb, nb * 512 * 2));
  DRL_CUDA(cudaMalloc(&e->w1q, kI8WBytes));
  DRL_CUDA(cudaMalloc(&e->c1scale, 32 * sizeof(float)));
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
  cudaFree(e->w2f); cudaFree(e->w3f); cudaFree(e->wff); cudaFree(e->wfd); cudaFree(e->w3d); cudaFree(e->w2d); cudaFree(e->whd); cudaFree(e->featb); cudaFree(e->w1q); cudaFree(e->c1scale); cudaFree(e->gt); cudaFree(e->bits[0]); cudaFree(e->bits[1]); cudaFree(e->bits[2]);
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
    a.job[nj] = PackArgs{w, out, mode, fp16, variants, KB, N, Cin, KH, KW, Cout, {nullptr, nullptr, nullptr, nullptr}};
    a.first_block[nj] = blocks;
    const int64_t total = (int64_t)variants * KB * N * 64;
    int b = (int)((total + 1023) / 1024);               // ~4 elements per thread
    blocks += b < 1 ? 1 : b;
    ++nj;
  };
  add(p + o[2], e->w2f, PK_CONV_FWD, 0, 1, 8, 64, 32, 4, 4, 64);
  add(p + o[4], e->w3f, PK_CONV_FWD, 0, 1, 9, 64, 64, 3, 3, 64);
  add(p + o[4], e->w3d, PK_CONV3_DGRAD, 0, 1, 9, 64, 64, 3, 3, 64);
  add(p + o[2], e->w2d, PK_CONV2_DGRAD, 0, 1, 4, 128, 32, 4, 4, 64);
  if (n->A + n->K <= kHmNH) {
    add(p + o[8], e->whd, PK_HEADS, 0, 1, 8, kHmNH, n->A, n->K, 1, 1);
    for (int k = 0; k < n->K; ++k) a.job[nj - 1].wv[k] = p + o[10 + 2 * k];
  }
  a.first_block[nj] = blocks;
  a.n_jobs = nj;
  a.wfc = p + o[6]; a.fc_fwd = (__nv_bfloat16*)e->wff; a.fc_dgrad = (__nv_bfloat16*)e->wfd; a.fc_blocks = sm_count() * 2;
  a.w1 = p + o[0]; a.scale = kInv255; a.w1q = e->w1q; a.c1scale = e->c1scale;
  pack_all_kernel<<<a.fc_blocks + blocks + 1, 256, 0, st>>>(a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// Kernel-variant switches (drl_debug_set; the measured defaults are the values below):
// [6] fc GEMMs: bits 0-1 CTA pairs (cta_group::2, 256 x 256 tiles): 0 none, 1 forward only (default), 2 forward + data
//     gradient; bit 2: the single-CTA data gradient with FOUR epilogue warpgroups (0.107 -> 0.102 ms) => default 5
// [7] descriptor-shifted conv kernels on CTA pairs (umma_tma_conv2.cuh), bit mask: 1 conv2 forward, 2 conv3 data
//     gradient, 4 conv2 data gradient.  Measured at 32 768 samples: conv2 fwd 0.255 -> 0.251 ms, conv3 dgrad 0.251 ->
//     0.232 ms, conv2 dgrad 0.253 -> 0.340 ms (it is HBM / epilogue bound: the pair only couples two epilogues) => 3.
//     bit 8: conv3 data gradient on CTA pairs with accumulator tiles that span samples (tma_conv2s_kernel, groups of 5
//     samples in 4 tiles): 0.235 -> 0.217 ms => default 11.
// [9] 1 = heads + losses on the CUDA-core kernel (heads_loss.cu) instead of the tensor-core kernel (umma_heads.cuh): 0.087 vs
//     ~0.02 ms at 32 768 samples => default 0
// [8] heads kernel compiled for 3 resident CTAs per SM (80 registers instead of 127, 64 bytes of spill) for the exact-(A, K)
//     training instantiations with A <= 8: 0.094 -> 0.084 ms => default 1
// (The cp.async row-gather / im2col engines of round 1 that flags [0]-[5] used to select were measured slower and are gone.)
int g_debug_flags[16] = {0, 0, 0, 0, 0, 0, 5, 11, 1, 0, 0, 0, 0, 0, 0, 0};

int bf16_trunk_forward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, bool feat_bf16, cudaStream_t st) {
  Bf16Engine* e = n->bf16;
  const float* p = n->params;
  const int64_t* o = n->off;
  // conv1 on the raw bytes: u8 x s8 integer MMAs over TMA-staged row phases of the frame, 16-bit fixed-point weights;
  // y = relu(acc * delta_c * (1/255) + b)   (scale_observations.py:35 fused; umma_conv1_i8.cuh)
  {
    if (reinterpret_cast<uintptr_t>(obs) & 15) return DRL_ERR_BAD_ARG;      // TMA: 16-byte aligned frame store
    ProfileScope prof("conv1_fwd", st);
    CUtensorMap tm;
    DRL_TRY(make_tmap_frame_phases(&tm, obs));
    Conv1I8Args a1{};
    a1.row_idx = row_idx; a1.nb = nb; a1.bpack = e->w1q; a1.cscale = e->c1scale; a1.bias = p + o[1];
    a1.out = reinterpret_cast<__nv_bfloat16*>(n->act[0]); a1.mask_bits = e->bits[0];
    static int configured[kMaxDevices] = {0};
    DRL_TRY(ensure_dyn_smem(conv1_fwd_i8_kernel, kI8Smem, configured));
    conv1_fwd_i8_kernel<<<nb < sm_count() ? nb : sm_count(), kI8Threads, kI8Smem, st>>>(tm, a1);
    DRL_LAUNCH_CHECK();
  }
  {
    // act1 seen as [nb, 20 rows, 10 pixel pairs, 64]; even / odd rows are staged as two 10x10 planes (TMA row
    // stride 2), so tap row ky reads plane ky&1 shifted by (ky>>1)*10 + half and the GEMM rows are oy*10 + ox
    TmaConvArgs c2 = tma_conv_args(8, 2, 10, 10, 0, 1, 9, 9, e->w2f, n->act[1], p + o[3], nullptr, 5184, 9 * 64, 64, 1, 0);
    c2.planes = 2; c2.bits_out = e->bits[1];
    DRL_TRY((launch_tma_conv<64, EPI_BIAS_RELU_BF16, 4, 4, 2, 128, 8, 2, 10, 2>("conv2_fwd", n->act[0], nb, 20, 10, 10, 20, c2, st, 1)));
  }
  {
    TmaConvArgs c3 = tma_conv_args(9, 3, 9, 9, 0, 1, 7, 7, e->w3f, n->act[2], p + o[5], nullptr, 3136, 7 * 64, 64, 1, 0);
    c3.bits_out = e->bits[2];
    DRL_TRY((launch_tma_conv<64, EPI_BIAS_RELU_BF16, 6, 4, 2, 64, 9, 3, 9, 1>("conv3_fwd", n->act[1], nb, 9, 9, 9, 9, c3, st)));
  }
  {
    TmaGemmArgs g{};
    g.M = nb; g.N_total = 512; g.KB = 49; g.out = n->feat; g.out_ld = 512; g.bias = p + o[7]; g.mask_src = nullptr;
    g.bias_grad = nullptr; g.bias_mod = 1; g.bits_in = nullptr;
    if (feat_bf16) {                                  // learner step: bf16 features for the tensor-core heads kernel
      g.out = e->featb;
      DRL_TRY((launch_tma_gemm<256, EPI_BIAS_RELU_BF16>("fc_fwd", n->act[2], nb, 3136, e->wff, 512, g, st)));
    } else {
      DRL_TRY((launch_tma_gemm<256, EPI_BIAS_RELU_F32>("fc_fwd", n->act[2], nb, 3136, e->wff, 512, g, st)));
    }
  }
  return DRL_OK;
}

// heads + losses (+ backward) on tensor cores from the bf16 features of the last bf16_trunk_forward(feat_bf16 = true)
bool bf16_heads_on_tensor_cores(const drl_net* n) { return n->A + n->K <= kHmNH && g_debug_flags[9] == 0; }

int bf16_heads_loss(drl_net* n, int mode, const HeadParams& hpp, const int64_t* row_idx, int nb, const drl_ppo_batch_t& batch,
                    const drl_ppo_hyper_t& hp, float* losses_out, cudaStream_t st) {
  return heads_mma_launch(mode, n->bf16->featb, n->bf16->whd, hpp, row_idx, nb, batch, hp, losses_out, n->dpre, st);
}

template <int MAXA, int MODE, int AX, int KX>
static int heads_mma_go(const CUtensorMap& tf, const HeadsMmaArgs& a, int nb, cudaStream_t st) {
  auto kern = heads_mma_kernel<MAXA, MODE, AX, KX>;
  static int configured[kMaxDevices] = {0};
  DRL_TRY(ensure_dyn_smem(kern, kHmSmem, configured));
  const int tiles = (nb + 127) / 128;
  kern<<<tiles < sm_count() ? tiles : sm_count(), kHmThreads, kHmSmem, st>>>(tf, a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

int heads_mma_launch(int mode, const void* featb, const void* wpack, const HeadParams& hpp, const int64_t* row_idx, int nb,
                     const drl_ppo_batch_t& batch, const drl_ppo_hyper_t& hp, float* losses_out, void* dpre, cudaStream_t st) {
  CUtensorMap tf;
  DRL_TRY(make_tmap_bf16(&tf, featb, (uint64_t)nb, 512, 128));
  HeadsMmaArgs a{};
  a.hp_ptrs = hpp; a.wpack = wpack; a.row_idx = row_idx; a.nb = nb; a.batch = batch; a.hp = hp; a.inv_nb = 1.0f / (float)nb;
  a.losses_out = losses_out; a.dpre = reinterpret_cast<__nv_bfloat16*>(dpre);
  const int A = hp.num_actions, K = hp.num_streams;
#define HM(MAXA, AX, KX)                                                      \
  do {                                                                        \
    if (mode == 2) return heads_mma_go<MAXA, 2, AX, KX>(tf, a, nb, st);       \
    return heads_mma_go<MAXA, 1, AX, KX>(tf, a, nb, st);                      \
  } while (0)
  if (A == 6 && K == 1) HM(8, 6, 1);          // Pong-class action sets (atari_ppo)
  if (A == 18 && K == 2) HM(18, 18, 2);       // RND + PPO (atari_rnd)
  if (A == 18 && K == 1) HM(18, 18, 1);
  if (A <= 8) HM(8, 0, 0);
  if (A <= 18) HM(18, 0, 0);
  HM(32, 0, 0);
#undef HM
}

int bf16_trunk_backward(drl_net* n, const uint8_t* obs, const int64_t* row_idx, int nb, float* g, bool overlap_comm,
                        cudaStream_t st) {
  Bf16Engine* e = n->bf16;
  const int64_t* o = n->off;
  const int sms = sm_count();

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
    w.gt = e->gt + kGtFc; w.gt_ld = 512;
    auto kern = wgrad_tma_kernel<256, 1>;
    static int configured[kMaxDevices] = {0};
    DRL_TRY(ensure_dyn_smem(kern, Cf::kSmem, configured));
    int gx = 2 * sms / 50;                       // two CTAs per SM
    if (gx > (nb + 31) / 32) gx = (nb + 31) / 32;
    if (gx < 1) gx = 1;
    kern<<<dim3(gx, 50), 192, Cf::kSmem, st>>>(tp, td, w);
    DRL_LAUNCH_CHECK();
  }
  DRL_TRY(wgrad_untranspose("fc_wgrad_t", e->gt + kGtFc, g + o[6], 64, 49, 512, st));
  // fc weights, fc bias and the heads (everything from the fc weight on = 95 % of the gradient) are final: start
  // their all-reduce on the side stream while the convolution gradients are computed.  Only the learner step asks for
  // this (drl_net_ppo_step joins the exchange afterwards); drl_net_backward never starts a collective.
  if (overlap_comm && n->comm && comm_world(n->comm) > 1) DRL_TRY(comm_allreduce_after(n->comm, g, o[6], n->nparams, st));
  {
    TmaGemmArgs a{};
    a.M = nb; a.N_total = 3136; a.KB = 8; a.out = n->dact[2]; a.out_ld = 3136; a.bias = nullptr;
    a.mask_src = (const __nv_bfloat16*)n->act[2];
    a.bias_grad = g + o[5]; a.bias_mod = 64;          // conv3 bias gradient = column sums of dact3 (column % 64)
    a.bits_in = e->bits[2];
    if (g_debug_flags[6] & 4) DRL_TRY((launch_tma_gemm<256, EPI_MASK_BF16, 4>("fc_dgrad", n->dpre, nb, 512, e->wfd, 3136, a, st)));   // 4 epilogue warpgroups
    else DRL_TRY((launch_tma_gemm<256, EPI_MASK_BF16>("fc_dgrad", n->dpre, nb, 512, e->wfd, 3136, a, st)));
  }
  // ---- conv3 (its bias gradient was accumulated by the fc data-gradient epilogue)
  {
    // taps t = (ty, tx) shift the staged 9x9 input by ty*9 + tx rows; blocks = tap pairs (0,1)(2,3)(4,5)(6,7)(8,-)
    WgradConvArgs w{};
    w.ksteps = 4; w.n_blocks = 5; w.p_planes = 1; w.p_pad = 0;
    const int sh[10] = {0, 1, 2, 9, 10, 11, 18, 19, 20, 21};
    for (int mb = 0; mb < 5; ++mb) { w.a_off[mb] = sh[2 * mb] * 128; w.a_lbo[mb] = (sh[2 * mb + 1] - sh[2 * mb]) * 128; }
    w.p_plane_bytes = 11264;                         // 88 rows >= 63 + 21 + 1
    w.gw = g + o[4]; w.Cin = 64; w.KH = 3; w.KW = 3; w.kp_valid = 576;
    w.gt = e->gt + kGtConv3;
    DRL_TRY((launch_wgrad_conv<64, 5, 4, 4>("conv3_wgrad", n->act[1], 9, 9, 9, 9, n->dact[2], 7, 7, 9, 8, nb, w, st)));
    DRL_TRY(wgrad_untranspose("conv3_wgrad_t", w.gt, g + o[4], 64, 9, 64, st));
  }
  {
    // dY3 zero-padded by 2 (TMA out-of-bounds fill) to 11x11; output rows y*11 + x over the 9x9 input grid
    TmaConvArgs c = tma_conv_args(9, 3, 11, 11, 2, 1, 9, 9, e->w3d, n->dact[1], nullptr, n->act[1], 5184, 9 * 64, 64, 1, 0);
    c.bias_grad = g + o[3]; c.bias_mod = 64;       // conv2 bias gradient
    c.bits_in = e->bits[1];
    if (g_debug_flags[7] & 8)   // groups of 5 samples on one row stream: 4 accumulator tiles instead of 5
      DRL_TRY((launch_tma_conv2s<64, EPI_MASK_BF16, 2, 2, 9, 3, 11, 5, 9, 4>("conv3_dgrad", n->dact[2], nb, 7, 7, 11, c, st)));
    else
      DRL_TRY((launch_tma_conv<64, EPI_MASK_BF16, 4, 4, 2, 128, 9, 3, 11, 1>("conv3_dgrad", n->dact[2], nb, 7, 7, 11, 11, c, st, 2)));
  }
  // ---- conv2
  {
    // act1 as pixel pairs, even / odd rows in two planes; block ky = atoms (ky, half 0/1): plane ky&1, shift (ky>>1)*10 + half
    WgradConvArgs w{};
    w.ksteps = 6; w.n_blocks = 4; w.p_planes = 2; w.p_pad = 0;
    w.p_plane_bytes = 14336;                         // 112 rows >= 95 + 11 + 1
    for (int ky = 0; ky < 4; ++ky) { w.a_off[ky] = (ky & 1) * w.p_plane_bytes + (ky >> 1) * 10 * 128; w.a_lbo[ky] = 128; }
    w.gw = g + o[2]; w.Cin = 32; w.KH = 4; w.KW = 4; w.kp_valid = 512;
    w.gt = e->gt + kGtConv2;
    DRL_TRY((launch_wgrad_conv<64, 4, 4, 6>("conv2_wgrad", n->act[0], 20, 10, 10, 20, n->dact[1], 9, 9, 10, 10, nb, w, st)));
    DRL_TRY(wgrad_untranspose("conv2_wgrad_t", w.gt, g + o[2], 32, 16, 64, st));
  }
  // conv2 + conv3 weights and biases are final (conv2's bias gradient came from conv3's data gradient): second overlapped
  // bucket; only conv1's 8 224 floats remain for the exchange after the last kernel
  if (overlap_comm && n->comm && comm_world(n->comm) > 1) DRL_TRY(comm_allreduce_after(n->comm, g, o[2], o[6], st));
  {
    // the four parity classes of the stride-2 transposed convolution share their 2x2-tap gather over dY2:
    // ONE GEMM with N = 4 classes x 32 ci; column chunk `cls` lands on output pixel (2u + cls/2, 2v + cls%2).
    // dY2 zero-padded by 1 to 11x11; GEMM rows u*11 + v over the 10x10 grid of 2x2 output blocks
    TmaConvArgs c = tma_conv_args(4, 2, 11, 11, 1, 1, 10, 10, e->w2d, n->dact[0], nullptr, n->act[0], 12800, 20 * 32, 32, 2, 1);
    c.bias_grad = g + o[1]; c.bias_mod = 32;       // conv1 bias gradient
    c.bits_in = e->bits[0];
    // FOUR epilogue warpgroups: with N = 128 (4 column chunks per row: mask, bf16 pack, 4 x 32 column sums) this layer
    // was bound by epilogue throughput — 2 / 3 / 4 warpgroups: 0.257 / 0.231 / 0.220 ms (the N = 64 layers do not gain)
    DRL_TRY((launch_tma_conv<128, EPI_MASK_BF16, 4, 4, 4, 128, 4, 2, 11, 1>("conv2_dgrad", n->dact[1], nb, 9, 9, 11, 11, c, st, 4)));
  }
  // ---- conv1 (no data gradient: observations need none, agent.py:66-67)
  {
    ProfileScope prof("conv1_wgrad", st);
    CUtensorMap tdy;
    DRL_TRY(make_tmap_dy1(&tdy, n->dact[0], (uint64_t)nb));
    Conv1S2dArgs a{};
    a.obs = obs; a.row_idx = row_idx; a.nb = nb; a.scale = kInv255; a.gw = g + o[0];
    static int configured[kMaxDevices] = {0};
    DRL_TRY(ensure_dyn_smem(conv1_wgrad_s2d_kernel, kS2dWgradSmem, configured));
    conv1_wgrad_s2d_kernel<<<nb < sms ? nb : sms, kS2dThreads, kS2dWgradSmem, st>>>(tdy, a);
    DRL_LAUNCH_CHECK();
  }
  return DRL_OK;
}

}  // namespace drl

// kernel-variant switches (see g_debug_flags above)
extern "C" int drl_debug_set(int key, int value) {
  if (key < 0 || key >= 16) return DRL_ERR_BAD_ARG;
  drl::g_debug_flags[key] = value;
  return DRL_OK;
}
