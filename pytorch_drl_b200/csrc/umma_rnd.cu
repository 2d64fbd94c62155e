// tcgen05 bf16 engine of the RND predictor (DRL_PRECISION_BF16): _TeacherNetwork / _StudentNetwork of
// drl/envs/wrappers/stateful/intrinsic/random_network_distillation.py:22-87, the predictor loss of `learn` (:225-232) and the
// intrinsic reward of `step` (:193-196).
//
// The RND trunk has 16 / 32 / 32 channels — half of NatureCNN's 32 / 64 / 64.  Instead of new operand layouts the engine packs
// SAMPLE PAIRS along the channel dimension: activations are [pairs][H][W][2 x C] (channel = s*C + c for sample 2p + s of pair p),
// weights are block-diagonal ([W 0; 0 W]).  conv2 then IS NatureCNN's conv2 geometry (32 -> 64, 4x4 / 2, 20x20 -> 9x9) and conv3 a
// 16-tap variant of its conv3, so forward, data-gradient and weight-gradient kernels are the descriptor-shifted TMA kernels of
// umma_tma_conv.cuh with LeakyReLU(0.2) epilogues.  The zero blocks double the MMA count, but these N = 64 MMAs are bound by the
// 128 B/clk operand fetch of the A tile, not by the tensor pipe, so a block-diagonal N = 64 step (48 cycles) costs about as
// much as the two N = 32 steps (2 x 40) it replaces — and both samples' outputs are useful.  In the weight gradients the
// off-diagonal blocks (products across the two samples) are discarded by the flush.  conv1 (1 input channel) has its own kernels
// (umma_rnd_conv1.cuh); the linear layers are plain TMA GEMMs (teacher 1152 -> 512; student 1152 -> 256 -> 256 -> 512).
#include "rnd.cuh"
#include <new>
#include "umma_launch.cuh"
#include "umma_rnd_conv1.cuh"

namespace drl {

constexpr int64_t kRGtFc1 = 0, kRGtConv3 = (int64_t)1152 * 256, kRGtConv2 = kRGtConv3 + 512 * 32, kRGtTotal = kRGtConv2 + 256 * 32;

struct RndBf16 {
  // packed operands
  void *w1c = nullptr;                               // conv1 [32 x 64] tile image (teacher 16 | student 16)
  void *tw2 = nullptr, *sw2 = nullptr;               // conv2 fwd [8][64 x 64] block-diagonal tile images
  void *tw3 = nullptr, *sw3 = nullptr;               // conv3 fwd [16][64 x 64]
  void *sw3d = nullptr, *sw2d = nullptr;             // student data-gradient images: conv3 [16][64 x 64] (flipped), conv2 [4][128 x 64]
  __nv_bfloat16 *twf = nullptr;                      // teacher fc [512][1152], k = pix*32 + c
  __nv_bfloat16 *sf1 = nullptr, *sf1d = nullptr;     // student fc1 [256][1152] and its transpose [1152][256]
  __nv_bfloat16 *sf2 = nullptr, *sf2d = nullptr;     // fc2 [256][256], transpose
  __nv_bfloat16 *sf3 = nullptr, *sf3d = nullptr;     // fc3 [512][256], transpose [256][512]
  float *b1c = nullptr, *tb2 = nullptr, *sb2 = nullptr, *tb3 = nullptr, *sb3 = nullptr;   // pair-duplicated conv biases
  float *na = nullptr, *nbias = nullptr;             // per-pixel normaliser scale / shift
  // activations (pair-packed trunk, per-sample from act3 on)
  __nv_bfloat16 *tact1 = nullptr, *sact1 = nullptr, *tact2 = nullptr, *sact2 = nullptr, *tact3 = nullptr, *sact3 = nullptr;
  uint32_t *sbits1 = nullptr, *sbits2 = nullptr, *sbits3 = nullptr;
  float *y = nullptr, *yh = nullptr;
  __nv_bfloat16 *h1 = nullptr, *h2 = nullptr, *dyh = nullptr, *dh2 = nullptr, *dh1 = nullptr;
  __nv_bfloat16 *dsact3 = nullptr, *dsact2 = nullptr, *dsact1 = nullptr;
  float* gt = nullptr;                               // weight-gradient scratch (zero between steps)
};

// ---- packing -------------------------------------------------------------------------------------------------------------
enum { RP_CONV1 = 0, RP_CONV2_FWD, RP_CONV3_FWD, RP_CONV3_DGRAD, RP_CONV2_DGRAD,      // swizzled tile images [KB][N x 64]
       RP_FC_PIX,        // [R][1152] bf16, k = pix*32 + c  <-  W[r][c*36 + pix]
       RP_FC_PIX_T,      // [1152][R]: transpose of the above
       RP_COPY,          // [R][C] plain
       RP_TRANSPOSE,     // [C][R] <- W[R][C]
       RP_BIAS_PAIR,     // fp32 [2][C] <- b[C] twice
       RP_BIAS_CONV1 };  // fp32 [32] <- (w[16] | w2[16])

struct RndPackJob { int mode; const float* w; const float* w2; void* out; int KB, N, R, C; int64_t total; };
struct RndPackArgs { RndPackJob job[20]; int first_block[21]; int n_jobs; };

__device__ __forceinline__ float rnd_pack_tile_src(const RndPackJob& j, int n, int k) {
  switch (j.mode) {
    case RP_CONV1: return n < 16 ? j.w[n * 64 + k] : j.w2[(n - 16) * 64 + k];
    case RP_CONV2_FWD: {           // n = (s', co32); k = (ky, kx, (s, ci16))
      const int ky = k >> 7, r2 = k & 127, kx = r2 >> 5, cip = r2 & 31, s = cip >> 4, ci = cip & 15, sp = n >> 5, co = n & 31;
      return s == sp ? j.w[((co * 16 + ci) * 4 + ky) * 4 + kx] : 0.f;
    }
    case RP_CONV3_FWD: {           // k = (tap = ky*4 + kx, (s, ci32))
      const int tap = k >> 6, cip = k & 63, s = cip >> 5, ci = cip & 31, sp = n >> 5, co = n & 31;
      return s == sp ? j.w[(co * 32 + ci) * 16 + tap] : 0.f;
    }
    case RP_CONV3_DGRAD: {         // B[n = (s, ci)][k = (ty, tx, (s', co))] = W3[co][ci][3 - ty][3 - tx]
      const int t = k >> 6, cop = k & 63, sp = cop >> 5, co = cop & 31, s = n >> 5, ci = n & 31, ty = t >> 2, tx = t & 3;
      return s == sp ? j.w[((co * 32 + ci) * 4 + (3 - ty)) * 4 + (3 - tx)] : 0.f;
    }
    default: {                     // RP_CONV2_DGRAD: n = (class (py, px), (s, ci16)); k = (ty, tx, (s', co32)): W2[co][ci][py + 2(1-ty)][px + 2(1-tx)]
      const int cls = n >> 5, cip = n & 31, s = cip >> 4, ci = cip & 15, py = cls >> 1, px = cls & 1;
      const int t = k >> 6, cop = k & 63, sp = cop >> 5, co = cop & 31, ty = t >> 1, tx = t & 1;
      return s == sp ? j.w[((co * 16 + ci) * 4 + (py + 2 * (1 - ty))) * 4 + (px + 2 * (1 - tx))] : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) rnd_pack_kernel(const RndPackArgs a) {
  int ji = 0;
  while (ji + 1 < a.n_jobs && (int)blockIdx.x >= a.first_block[ji + 1]) ++ji;
  const RndPackJob& j = a.job[ji];
  const int nblk = a.first_block[ji + 1] - a.first_block[ji], lb = blockIdx.x - a.first_block[ji];
  for (int64_t e = (int64_t)lb * blockDim.x + threadIdx.x; e < j.total; e += (int64_t)nblk * blockDim.x) {
    if (j.mode <= RP_CONV2_DGRAD) {
      const int kk = (int)(e & 63), n = (int)((e >> 6) % j.N), kb = (int)((e >> 6) / j.N);
      const float v = rnd_pack_tile_src(j, n, kb * 64 + kk);
      const int64_t off = (int64_t)kb * j.N * 128 + umma::swz_off<128>((uint32_t)n, (uint32_t)(kk >> 3)) + (kk & 7) * 2;
      *reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(j.out) + off) = __float2bfloat16_rn(v);
    } else if (j.mode == RP_FC_PIX) {
      const int k = (int)(e % 1152), r = (int)(e / 1152), pix = k >> 5, c = k & 31;
      reinterpret_cast<__nv_bfloat16*>(j.out)[e] = __float2bfloat16_rn(j.w[(int64_t)r * 1152 + c * 36 + pix]);
    } else if (j.mode == RP_FC_PIX_T) {
      const int r = (int)(e % j.R), k = (int)(e / j.R), pix = k >> 5, c = k & 31;
      reinterpret_cast<__nv_bfloat16*>(j.out)[e] = __float2bfloat16_rn(j.w[(int64_t)r * 1152 + c * 36 + pix]);
    } else if (j.mode == RP_COPY) {
      reinterpret_cast<__nv_bfloat16*>(j.out)[e] = __float2bfloat16_rn(j.w[e]);
    } else if (j.mode == RP_TRANSPOSE) {
      const int r = (int)(e % j.R), c = (int)(e / j.R);
      reinterpret_cast<__nv_bfloat16*>(j.out)[e] = __float2bfloat16_rn(j.w[(int64_t)r * j.C + c]);
    } else if (j.mode == RP_BIAS_PAIR) {
      reinterpret_cast<float*>(j.out)[e] = j.w[e % j.C];
    } else {
      reinterpret_cast<float*>(j.out)[e] = e < 16 ? j.w[e] : j.w2[e - 16];
    }
  }
}

int rnd_bf16_pack(drl_rnd* r, bool teacher, bool student, cudaStream_t st) {
  ProfileScope prof("rnd_pack", st);
  RndBf16* e = r->bf16;
  const float *tp = r->tp, *sp = r->sp;
  const int64_t *to = r->toff, *so = r->soff;
  RndPackArgs a{};
  int nj = 0, blocks = 0;
  auto add = [&](int mode, const float* w, const float* w2, void* out, int KB, int N, int R, int C, int64_t total) {
    a.job[nj] = RndPackJob{mode, w, w2, out, KB, N, R, C, total};
    a.first_block[nj] = blocks;
    int b = (int)((total + 2047) / 2048);
    blocks += b < 1 ? 1 : (b > 96 ? 96 : b);
    ++nj;
  };
  auto tile = [&](int mode, const float* w, void* out, int KB, int N) { add(mode, w, nullptr, out, KB, N, 0, 0, (int64_t)KB * N * 64); };
  // conv1 holds both networks: repacked whenever either changes
  add(RP_CONV1, tp + to[0], sp + so[0], e->w1c, 1, 32, 0, 0, 32 * 64);
  add(RP_BIAS_CONV1, tp + to[1], sp + so[1], e->b1c, 0, 0, 0, 0, 32);
  if (teacher) {
    tile(RP_CONV2_FWD, tp + to[2], e->tw2, 8, 64);
    tile(RP_CONV3_FWD, tp + to[4], e->tw3, 16, 64);
    add(RP_BIAS_PAIR, tp + to[3], nullptr, e->tb2, 0, 0, 0, 32, 64);
    add(RP_BIAS_PAIR, tp + to[5], nullptr, e->tb3, 0, 0, 0, 32, 64);
    add(RP_FC_PIX, tp + to[6], nullptr, e->twf, 0, 0, 512, 1152, (int64_t)512 * 1152);
  }
  if (student) {
    tile(RP_CONV2_FWD, sp + so[2], e->sw2, 8, 64);
    tile(RP_CONV3_FWD, sp + so[4], e->sw3, 16, 64);
    tile(RP_CONV3_DGRAD, sp + so[4], e->sw3d, 16, 64);
    tile(RP_CONV2_DGRAD, sp + so[2], e->sw2d, 4, 128);
    add(RP_BIAS_PAIR, sp + so[3], nullptr, e->sb2, 0, 0, 0, 32, 64);
    add(RP_BIAS_PAIR, sp + so[5], nullptr, e->sb3, 0, 0, 0, 32, 64);
    add(RP_FC_PIX, sp + so[6], nullptr, e->sf1, 0, 0, 256, 1152, (int64_t)256 * 1152);
    add(RP_FC_PIX_T, sp + so[6], nullptr, e->sf1d, 0, 0, 256, 1152, (int64_t)256 * 1152);
    add(RP_COPY, sp + so[8], nullptr, e->sf2, 0, 0, 256, 256, 256 * 256);
    add(RP_TRANSPOSE, sp + so[8], nullptr, e->sf2d, 0, 0, 256, 256, 256 * 256);
    add(RP_COPY, sp + so[10], nullptr, e->sf3, 0, 0, 512, 256, 512 * 256);
    add(RP_TRANSPOSE, sp + so[10], nullptr, e->sf3d, 0, 0, 512, 256, 512 * 256);
  }
  a.first_block[nj] = blocks;
  a.n_jobs = nj;
  rnd_pack_kernel<<<blocks, 256, 0, st>>>(a);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// ---- allocation ------------------------------------------------------------------------------------------------------------
int rnd_bf16_alloc(drl_rnd* r) {
  RndBf16* e = new (std::nothrow) RndBf16();
  if (!e) return DRL_ERR_CUDA;
  r->bf16 = e;
  const size_t nb = (size_t)r->max_batch, np = (nb + 1) / 2;
  auto al = [](auto** p, size_t bytes) { return cudaMalloc(reinterpret_cast<void**>(p), bytes); };
  DRL_CUDA(al(&e->w1c, 4096));
  DRL_CUDA(al(&e->tw2, 8 * 8192)); DRL_CUDA(al(&e->sw2, 8 * 8192));
  DRL_CUDA(al(&e->tw3, 16 * 8192)); DRL_CUDA(al(&e->sw3, 16 * 8192));
  DRL_CUDA(al(&e->sw3d, 16 * 8192)); DRL_CUDA(al(&e->sw2d, 4 * 16384));
  DRL_CUDA(al(&e->twf, (size_t)512 * 1152 * 2));
  DRL_CUDA(al(&e->sf1, (size_t)256 * 1152 * 2)); DRL_CUDA(al(&e->sf1d, (size_t)256 * 1152 * 2));
  DRL_CUDA(al(&e->sf2, 256 * 256 * 2)); DRL_CUDA(al(&e->sf2d, 256 * 256 * 2));
  DRL_CUDA(al(&e->sf3, 512 * 256 * 2)); DRL_CUDA(al(&e->sf3d, 512 * 256 * 2));
  DRL_CUDA(al(&e->b1c, 32 * 4));
  DRL_CUDA(al(&e->tb2, 64 * 4)); DRL_CUDA(al(&e->sb2, 64 * 4)); DRL_CUDA(al(&e->tb3, 64 * 4)); DRL_CUDA(al(&e->sb3, 64 * 4));
  DRL_CUDA(al(&e->na, 7056 * 4)); DRL_CUDA(al(&e->nbias, 7056 * 4));
  DRL_CUDA(al(&e->tact1, np * 12800 * 2)); DRL_CUDA(al(&e->sact1, np * 12800 * 2));
  DRL_CUDA(al(&e->tact2, np * 5184 * 2)); DRL_CUDA(al(&e->sact2, np * 5184 * 2));
  DRL_CUDA(al(&e->tact3, np * 2304 * 2)); DRL_CUDA(al(&e->sact3, np * 2304 * 2));
  DRL_CUDA(al(&e->sbits1, np * 400 * 4)); DRL_CUDA(al(&e->sbits2, np * 162 * 4)); DRL_CUDA(al(&e->sbits3, np * 72 * 4));
  DRL_CUDA(al(&e->y, nb * 512 * 4)); DRL_CUDA(al(&e->yh, nb * 512 * 4));
  DRL_CUDA(al(&e->h1, nb * 256 * 2)); DRL_CUDA(al(&e->h2, nb * 256 * 2));
  DRL_CUDA(al(&e->dyh, nb * 512 * 2)); DRL_CUDA(al(&e->dh2, nb * 256 * 2)); DRL_CUDA(al(&e->dh1, nb * 256 * 2));
  DRL_CUDA(al(&e->dsact3, np * 2304 * 2)); DRL_CUDA(al(&e->dsact2, np * 5184 * 2)); DRL_CUDA(al(&e->dsact1, np * 12800 * 2));
  DRL_CUDA(al(&e->gt, kRGtTotal * 4));
  DRL_CUDA(cudaMemset(e->gt, 0, kRGtTotal * 4));
  return DRL_OK;
}

void rnd_bf16_free(drl_rnd* r) {
  RndBf16* e = r->bf16;
  if (!e) return;
  void* all[] = {e->w1c, e->tw2, e->sw2, e->tw3, e->sw3, e->sw3d, e->sw2d, e->twf, e->sf1, e->sf1d, e->sf2, e->sf2d, e->sf3, e->sf3d,
                 e->b1c, e->tb2, e->sb2, e->tb3, e->sb3, e->na, e->nbias, e->tact1, e->sact1, e->tact2, e->sact2, e->tact3, e->sact3,
                 e->sbits1, e->sbits2, e->sbits3, e->y, e->yh, e->h1, e->h2, e->dyh, e->dh2, e->dh1, e->dsact3, e->dsact2, e->dsact1,
                 e->gt};
  for (void* p : all) cudaFree(p);
  delete e;
  r->bf16 = nullptr;
}

// ---- predictor error -------------------------------------------------------------------------------------------------------
// err[i] = mean_d (y - yh)^2; loss += sum_i err[i] / nb; dyh = 2 (yh - y) / (D nb) as bf16 (the backward GEMMs' operand);
// gb3 += column sums of dyh (the bias gradient of the student's last linear layer).  One warp per sample.
__global__ void __launch_bounds__(256) rnd_mse_bf16_kernel(const float* __restrict__ y, const float* __restrict__ yh, int nb,
                                                           float* __restrict__ err, float* __restrict__ loss,
                                                           __nv_bfloat16* __restrict__ dyh, float* __restrict__ gb3) {
  constexpr int D = 512;
  const int lane = threadIdx.x & 31, wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const float gs = 2.f / ((float)D * (float)nb);
  float lacc = 0.f;
  float cs[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) cs[q] = 0.f;
  for (int i = wid; i < nb; i += nw) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {                      // lane owns columns lane*4 + 128*q .. +3
      const size_t e = (size_t)i * D + lane * 4 + 128 * q;
      const float4 a = *reinterpret_cast<const float4*>(y + e), b = *reinterpret_cast<const float4*>(yh + e);
      const float d0 = b.x - a.x, d1 = b.y - a.y, d2 = b.z - a.z, d3 = b.w - a.w;
      s = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s))));
      if (dyh) {
        const float g0 = gs * d0, g1 = gs * d1, g2 = gs * d2, g3 = gs * d3;
        *reinterpret_cast<uint2*>(dyh + e) = make_uint2(umma::pack_bf16x2(g0, g1), umma::pack_bf16x2(g2, g3));
        cs[4 * q] += g0; cs[4 * q + 1] += g1; cs[4 * q + 2] += g2; cs[4 * q + 3] += g3;
      }
    }
    s = warp_sum(s) / (float)D;
    if (lane == 0) {
      if (err) err[i] = s;
      lacc += s;
    }
  }
  if (loss && lane == 0 && lacc != 0.f) atomicAdd(loss, lacc / (float)nb);
  if (gb3) {                                           // block-level reduction first: one atomic per column and block
    __shared__ float red[8][512];
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<float4*>(&red[w][lane * 4 + 128 * q]) = make_float4(cs[4 * q], cs[4 * q + 1], cs[4 * q + 2], cs[4 * q + 3]);
    __syncthreads();
    for (int c = threadIdx.x; c < 512; c += 256) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][c];
      atomicAdd(gb3 + c, s);
    }
  }
}

// column sums of a bf16 [R, C] matrix (C = 2 * blockDim.x): bias gradients of the student's hidden layers
__global__ void __launch_bounds__(256) rnd_colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t R, int C,
                                                              float* __restrict__ out) {
  float a0 = 0.f, a1 = 0.f;
  for (int64_t r = blockIdx.x; r < R; r += gridDim.x) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(x + r * C + threadIdx.x * 2);
    a0 += __uint_as_float(w << 16);
    a1 += __uint_as_float(w & 0xffff0000u);
  }
  atomicAdd(out + threadIdx.x * 2, a0);
  atomicAdd(out + threadIdx.x * 2 + 1, a1);
}

static int rnd_colsum(const char* site, const __nv_bfloat16* x, int64_t R, int C, float* out, cudaStream_t st) {
  ProfileScope prof(site, st);
  int blocks = (int)(R < (int64_t)sm_count() * 4 ? R : (int64_t)sm_count() * 4);
  rnd_colsum_bf16_kernel<<<blocks, C / 2, 0, st>>>(x, R, C, out);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// ---- tensor maps ---------------------------------------------------------------------------------------------------------
// pair-packed dY1 [pairs][400][32] bf16 -> boxes of (32, 128 rows, 1), 64B swizzle, rows >= 400 zero fill
static int make_tmap_rnd_dy1(CUtensorMap* m, const void* base, uint64_t np) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return DRL_ERR_CUDA;
  const cuuint64_t dims[3] = {32, 400, np};
  const cuuint64_t strides[2] = {64, 400 * 64};
  const cuuint32_t box[3] = {32, 128, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DRL_OK : DRL_ERR_CUDA;
}

static RndConv1Args conv1_args(drl_rnd* r, const uint8_t* obs, const int64_t* row_idx, int nb) {
  RndBf16* e = r->bf16;
  RndConv1Args a{};
  a.obs = obs; a.row_idx = row_idx; a.nb = nb; a.na = e->na; a.nbias = e->nbias; a.bpack = e->w1c; a.bias = e->b1c;
  a.tout = e->tact1; a.sout = e->sact1; a.sbits = e->sbits1; a.gw = nullptr;
  return a;
}

static int fc_wgrad(const char* site, const void* p, int K, const void* dy, int N, int nb, float* gw, int Cin, int KH, int KW, float* gt,
                    cudaStream_t st) {
  ProfileScope prof(site, st);
  using Cf = WgradTmaCfg<256, 1>;
  CUtensorMap tp, td;
  DRL_TRY(make_tmap_bf16(&tp, p, (uint64_t)nb, (uint64_t)K, 32));
  DRL_TRY(make_tmap_bf16(&td, dy, (uint64_t)nb, (uint64_t)N, 32));
  WgradTmaArgs w{};
  w.M = nb; w.n_tiles = N / 256; w.kp_valid = K; w.gw = gw; w.Cin = Cin; w.KH = KH; w.KW = KW; w.gt = gt; w.gt_ld = N;
  auto kern = wgrad_tma_kernel<256, 1>;
  static int configured[kMaxDevices] = {0};
  DRL_TRY(ensure_dyn_smem(kern, Cf::kSmem, configured));
  const int gy = w.n_tiles * ((K + 127) / 128);
  int gx = 2 * sm_count() / gy;
  if (gx > (nb + 31) / 32) gx = (nb + 31) / 32;
  if (gx < 1) gx = 1;
  kern<<<dim3(gx, gy), 192, Cf::kSmem, st>>>(tp, td, w);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// ---- forward ---------------------------------------------------------------------------------------------------------------
int rnd_bf16_forward(drl_rnd* r, const uint8_t* obs, const int64_t* row_idx, int nb, const float* m1, const float* m2,
                     bool keep, cudaStream_t st) {
  RndBf16* e = r->bf16;
  const float *tp = r->tp, *sp = r->sp;
  const int64_t *to = r->toff, *so = r->soff;
  const int np = (nb + 1) / 2;
  {
    ProfileScope prof("rnd_norm_table", st);
    rnd_norm_table_kernel<<<(7056 + 255) / 256, 256, 0, st>>>(m1, m2, e->na, e->nbias, 7056);
    DRL_LAUNCH_CHECK();
  }
  {
    ProfileScope prof("rnd_conv1_fwd", st);
    RndConv1Args a = conv1_args(r, obs, row_idx, nb);
    if (!keep) a.sbits = nullptr;
    static int configured[kMaxDevices] = {0};
    DRL_TRY(ensure_dyn_smem(rnd_conv1_fwd_kernel, kR1FwdSmem, configured));
    rnd_conv1_fwd_kernel<<<np < sm_count() ? np : sm_count(), kR1FThreads, kR1FwdSmem, st>>>(a);
    DRL_LAUNCH_CHECK();
  }
  for (int net = 0; net < 2; ++net) {                // 0 teacher, 1 student
    const bool s = net == 1;
    // conv2: act1 pairs as [np, 20 rows, 10 pixel pairs, 64]; same staging as NatureCNN's conv2 (umma_net.cu)
    TmaConvArgs c2 = tma_conv_args(8, 2, 10, 10, 0, 1, 9, 9, s ? e->sw2 : e->tw2, s ? e->sact2 : e->tact2, s ? e->sb2 : e->tb2,
                                   nullptr, 5184, 9 * 64, 64, 1, 0);
    c2.planes = 2; c2.bits_out = (s && keep) ? e->sbits2 : nullptr;
    DRL_TRY((launch_tma_conv<64, EPI_BIAS_LEAKY_BF16, 4, 4, 2, 128, 8, 2, 10, 2>(s ? "rnd_conv2_fwd_s" : "rnd_conv2_fwd_t",
                                                                                s ? e->sact1 : e->tact1, np, 20, 10, 10, 20, c2, st)));
    // conv3: 4x4 / 1 over 9x9 -> 6x6: 16 taps shifting the staged image by ty*9 + tx rows, GEMM rows y*9 + x (54 <= 64);
    // the two samples of a pair (column chunks 0 / 1) go to their own per-sample [36 pixels][32] rows (k = pix*32 + c)
    TmaConvArgs c3 = tma_conv_args(16, 4, 9, 9, 0, 1, 6, 6, s ? e->sw3 : e->tw3, s ? e->sact3 : e->tact3, s ? e->sb3 : e->tb3,
                                   nullptr, 2304, 6 * 32, 32, 1, 0);
    c3.chunk_pitch = 1152; c3.bits_out = (s && keep) ? e->sbits3 : nullptr;
    DRL_TRY((launch_tma_conv<64, EPI_BIAS_LEAKY_BF16, 6, 4, 2, 64, 16, 4, 9, 1>(s ? "rnd_conv3_fwd_s" : "rnd_conv3_fwd_t",
                                                                               s ? e->sact2 : e->tact2, np, 9, 9, 9, 9, c3, st)));
  }
  auto gemm_args = [&](int N, int K, void* out, const float* bias) {
    TmaGemmArgs g{};
    g.M = nb; g.N_total = N; g.KB = K / 64; g.out = out; g.out_ld = N; g.bias = bias; g.bias_mod = 1;
    return g;
  };
  DRL_TRY((launch_tma_gemm<256, EPI_BIAS_F32>("rnd_fc_t", e->tact3, nb, 1152, e->twf, 512, gemm_args(512, 1152, e->y, tp + to[7]), st)));
  DRL_TRY((launch_tma_gemm<256, EPI_BIAS_RELU_BF16>("rnd_fc1", e->sact3, nb, 1152, e->sf1, 256, gemm_args(256, 1152, e->h1, sp + so[7]), st)));
  DRL_TRY((launch_tma_gemm<256, EPI_BIAS_RELU_BF16>("rnd_fc2", e->h1, nb, 256, e->sf2, 256, gemm_args(256, 256, e->h2, sp + so[9]), st)));
  DRL_TRY((launch_tma_gemm<256, EPI_BIAS_F32>("rnd_fc3", e->h2, nb, 256, e->sf3, 512, gemm_args(512, 256, e->yh, sp + so[11]), st)));
  return DRL_OK;
}

int rnd_bf16_mse(drl_rnd* r, int nb, float* err_out, float* loss_out, bool with_grad, float* grads, cudaStream_t st) {
  ProfileScope prof("rnd_mse", st);
  RndBf16* e = r->bf16;
  int blocks = (nb + 7) / 8;
  if (blocks > sm_count() * 4) blocks = sm_count() * 4;
  rnd_mse_bf16_kernel<<<blocks, 256, 0, st>>>(e->y, e->yh, nb, err_out, loss_out, with_grad ? e->dyh : nullptr,
                                             with_grad ? grads + r->soff[11] : nullptr);
  DRL_LAUNCH_CHECK();
  return DRL_OK;
}

// ---- student backward (random_network_distillation.py:232 loss.backward()) -------------------------------------------------
int rnd_bf16_backward(drl_rnd* r, const uint8_t* obs, const int64_t* row_idx, int nb, float* g, cudaStream_t st) {
  RndBf16* e = r->bf16;
  const int64_t* so = r->soff;
  const int np = (nb + 1) / 2;
  auto dgrad_args = [&](int N, int K, void* out, const __nv_bfloat16* mask_src) {
    TmaGemmArgs a{};
    a.M = nb; a.N_total = N; a.KB = K / 64; a.out = out; a.out_ld = N; a.mask_src = mask_src; a.bias_mod = 64;
    return a;
  };
  // fc3 (512 <- 256): bias gradient came from the MSE kernel
  DRL_TRY(fc_wgrad("rnd_fc3_wgrad", e->h2, 256, e->dyh, 512, nb, g + so[10], 256, 1, 1, nullptr, st));
  DRL_TRY((launch_tma_gemm<256, EPI_MASK_BF16>("rnd_fc3_dgrad", e->dyh, nb, 512, e->sf3d, 256, dgrad_args(256, 512, e->dh2, e->h2), st)));
  DRL_TRY(rnd_colsum("rnd_fc2_bgrad", e->dh2, nb, 256, g + so[9], st));
  // fc2
  DRL_TRY(fc_wgrad("rnd_fc2_wgrad", e->h1, 256, e->dh2, 256, nb, g + so[8], 256, 1, 1, nullptr, st));
  DRL_TRY((launch_tma_gemm<256, EPI_MASK_BF16>("rnd_fc2_dgrad", e->dh2, nb, 256, e->sf2d, 256, dgrad_args(256, 256, e->dh1, e->h1), st)));
  DRL_TRY(rnd_colsum("rnd_fc1_bgrad", e->dh1, nb, 256, g + so[7], st));
  // fc1 (256 <- 1152 = 36 pixels x 32 channels): scratch rows kp = pix*32 + c, untransposed to [256][c*36 + pix]
  DRL_TRY(fc_wgrad("rnd_fc1_wgrad", e->sact3, 1152, e->dh1, 256, nb, g + so[6], 32, 6, 6, e->gt + kRGtFc1, st));
  DRL_TRY(wgrad_untranspose("rnd_fc1_wgrad_t", e->gt + kRGtFc1, g + so[6], 32, 36, 256, st));
  {
    // data gradient into the PAIR-PACKED [np][36][2 x 32] map conv3's backward kernels read; LeakyReLU mask from act3's sign bits;
    // conv3 bias gradient = column sums folded over the 36 pixels (column % 32)
    TmaGemmArgs a = dgrad_args(1152, 256, e->dsact3, nullptr);
    a.bits_in = e->sbits3; a.pair_out = 1; a.bias_grad = g + so[5]; a.bias_mod = 32;
    DRL_TRY((launch_tma_gemm<256, EPI_MASK_LEAKY_BF16>("rnd_fc1_dgrad", e->dh1, nb, 256, e->sf1d, 1152, a, st)));
  }
  // conv3: weight gradient (8 accumulator blocks = 16 taps x 64 pair channels; diagonal blocks folded by the flush)
  {
    WgradConvArgs w{};
    w.ksteps = 4; w.n_blocks = 8; w.p_planes = 1; w.p_pad = 0;
    for (int ty = 0; ty < 4; ++ty)
      for (int j = 0; j < 2; ++j) { w.a_off[ty * 2 + j] = (ty * 9 + 2 * j) * 128; w.a_lbo[ty * 2 + j] = 128; }
    w.p_plane_bytes = 12288;                         // 96 rows >= 63 + 30 + 1
    w.gw = g + so[4]; w.Cin = 64; w.KH = 4; w.KW = 4; w.kp_valid = 1024;
    w.gt = e->gt + kRGtConv3; w.pair_c = 32;
    DRL_TRY((launch_wgrad_conv<64, 8, 4, 4>("rnd_conv3_wgrad", e->sact2, 9, 9, 9, 9, e->dsact3, 6, 6, 9, 8, np, w, st)));
    DRL_TRY(wgrad_untranspose("rnd_conv3_wgrad_t", w.gt, g + so[4], 32, 16, 32, st));
  }
  {
    // dY3 zero-padded by 3 (TMA out-of-bounds fill) to 12x12; output rows y*12 + x over the 9x9 input grid (108 <= 128)
    TmaConvArgs c = tma_conv_args(16, 4, 12, 12, 3, 1, 9, 9, e->sw3d, e->dsact2, nullptr, nullptr, 5184, 9 * 64, 64, 1, 0);
    c.bias_grad = g + so[3]; c.bias_mod = 32; c.bits_in = e->sbits2;
    DRL_TRY((launch_tma_conv<64, EPI_MASK_LEAKY_BF16, 4, 4, 2, 128, 16, 4, 12, 1>("rnd_conv3_dgrad", e->dsact3, np, 6, 6, 12, 12, c, st)));
  }
  // conv2: NatureCNN's conv2 geometry on the pair-packed maps
  {
    WgradConvArgs w{};
    w.ksteps = 6; w.n_blocks = 4; w.p_planes = 2; w.p_pad = 0;
    w.p_plane_bytes = 14336;
    for (int ky = 0; ky < 4; ++ky) { w.a_off[ky] = (ky & 1) * w.p_plane_bytes + (ky >> 1) * 10 * 128; w.a_lbo[ky] = 128; }
    w.gw = g + so[2]; w.Cin = 32; w.KH = 4; w.KW = 4; w.kp_valid = 512;
    w.gt = e->gt + kRGtConv2; w.pair_c = 16;
    DRL_TRY((launch_wgrad_conv<64, 4, 4, 6>("rnd_conv2_wgrad", e->sact1, 20, 10, 10, 20, e->dsact2, 9, 9, 10, 10, np, w, st)));
    DRL_TRY(wgrad_untranspose("rnd_conv2_wgrad_t", w.gt, g + so[2], 16, 16, 32, st));
  }
  {
    TmaConvArgs c = tma_conv_args(4, 2, 11, 11, 1, 1, 10, 10, e->sw2d, e->dsact1, nullptr, nullptr, 12800, 20 * 32, 32, 2, 1);
    c.bias_grad = g + so[1]; c.bias_mod = 16; c.bits_in = e->sbits1;
    DRL_TRY((launch_tma_conv<128, EPI_MASK_LEAKY_BF16, 4, 4, 4, 128, 4, 2, 11, 1>("rnd_conv2_dgrad", e->dsact2, np, 9, 9, 11, 11, c, st)));
  }
  // conv1 weight gradient
  {
    ProfileScope prof("rnd_conv1_wgrad", st);
    CUtensorMap tdy;
    DRL_TRY(make_tmap_rnd_dy1(&tdy, e->dsact1, (uint64_t)np));
    RndConv1Args a = conv1_args(r, obs, row_idx, nb);
    a.gw = g + so[0];
    static int configured[kMaxDevices] = {0};
    DRL_TRY(ensure_dyn_smem(rnd_conv1_wgrad_kernel, kR1WgradSmem, configured));
    rnd_conv1_wgrad_kernel<<<np < sm_count() ? np : sm_count(), kR1Threads, kR1WgradSmem, st>>>(tdy, a);
    DRL_LAUNCH_CHECK();
  }
  return DRL_OK;
}

}  // namespace drl
