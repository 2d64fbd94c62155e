// tcgen05 implicit-GEMM kernels of the bf16 engine.
//
// Family 1 — `rowgather_gemm_kernel`:  D[128 x N] = A[128 x K] * B[N x K]^T, both operands K-major
// in 128B-swizzled shared-memory tiles.
//   * A rows are gathered on the fly (im2col) by 4 producer warps: output row m = (sample, y, x)
//     of an output grid, k-block kb = (tap row, 64-element run inside it).  conv1 reads uint8
//     frames (through the minibatch row index) and converts to fp16 exactly; the other layers read
//     bf16 NHWC activations.  Out-of-range taps (padding of the data-gradient convolutions) are
//     zero-filled.
//   * B tiles are weights pre-packed on the device into ready-made swizzled [N x 64] tile images
//     (umma_pack.cu) and land with one 1-D TMA bulk copy per k-block (UBLKCP), completion on the
//     stage's mbarrier.
//   * one elected thread issues tcgen05.mma (M=128, N, K=16) into a double-buffered TMEM
//     accumulator; tcgen05.commit releases the smem stage / publishes the accumulator.
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b) -> bias/scale/ReLU or ReLU-mask -> global.
// Persistent CTAs: blockIdx.x strides over M tiles, blockIdx.y selects the N tile / weight variant.
//
// Family 2 — `wgrad_gemm_kernel`:  dW^T[Kp x N] += P[m x Kp]^T * dY[m x N], reduction over the
// rows m; both operands MN-major (the tiles are stored exactly as they lie in memory: one row per
// m, 128 or 64 bytes of contiguous kp / channels), accumulators for all of Kp stay in TMEM, the
// partial results of the CTAs (split over m) are added to the fp32 gradient with red.global.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace drl {

// ---- family 1 -------------------------------------------------------------------------------------
enum { SRC_BF16 = 0, SRC_U8 = 1 };
enum { EPI_BIAS_RELU_BF16 = 0, EPI_BIAS_RELU_F32 = 1, EPI_MASK_BF16 = 2 };

struct RowGatherArgs {
  // problem
  int M;                 // output rows = nb * P
  int KB;                // number of 64-element k-blocks
  // A gather
  const void* src;       // uint8 [rows,84,84,4] or bf16 NHWC [nb,H,W,C]
  const int64_t* row_idx;  // sample -> source row (conv1 minibatch gather) or nullptr
  int P, OW;             // output grid positions per sample, grid width
  int H, W, C;           // source map
  int stride, pad;       // source (y,x) = (oy*stride - pad + ty, ox*stride - pad + tx)
  int KBX;               // k-blocks per tap row (KW*C/64); SRC_U8: unused
  // B
  const void* bpack;     // [variants][KB][N x 64] swizzled tile images
  int64_t b_variant_stride;   // bytes between variants (blockIdx.y)
  // epilogue
  void* out;             // bf16 or fp32
  const float* bias;     // EPI_BIAS_*: [N_total]
  float scale;           // acc * scale + bias
  const __nv_bfloat16* mask_src;   // EPI_MASK: activation with the layout of `out`
  int64_t out_sample_pitch;   // elements
  int out_row_pitch;          // elements per output map row  (out W * out C)
  int out_C;                  // elements per output pixel
  int osy, osx;               // output pixel = (oy*osy + ooy, ox*osx + oox)
  int n_variant_stride;       // n0 = blockIdx.y * n_variant_stride (N tiles); 0 for parity classes
  int classes;                // 1: 32-column chunk ch of the tile goes to output pixel (oy*osy + ch/2, ox*osx + ch%2)
};

template <int N>
struct RowGatherSmem {
  static constexpr int kStages = N >= 128 ? 3 : 4;     // 96 KB per CTA either way: two CTAs per SM
  static constexpr int kABytes = 128 * 128;
  static constexpr int kBBytes = N * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256 + 1024;   // barriers + bias[N]
  static constexpr int kTotal = kStages * kStageBytes + kBarBytes + 1024;   // + alignment slack
};

template <int N, int SRC, int EPI>
__global__ void __launch_bounds__(288, 2) rowgather_gemm_kernel(const RowGatherArgs args) {
  using namespace umma;
  using L = RowGatherSmem<N>;
  constexpr int S = L::kStages;
  constexpr int ACC_COLS = 2 * N;                       // double-buffered accumulator
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // S x 16 KB
  uint8_t* sB = smem + S * L::kABytes;                  // S x N*128
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * L::kStageBytes);
  uint64_t* full = bars;             // [S]
  uint64_t* empty = bars + S;        // [S]
  uint64_t* tfull = bars + 2 * S;    // [2]
  uint64_t* tempty = bars + 2 * S + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  float* sBias = reinterpret_cast<float*>(bars) + 64;   // [N] (256 B past the barriers)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (args.M + 127) / 128;
  if (EPI != EPI_MASK_BF16 && threadIdx.x < N) sBias[threadIdx.x] = args.bias[blockIdx.y * args.n_variant_stride + threadIdx.x];

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 129); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint8_t* bpack = reinterpret_cast<const uint8_t*>(args.bpack) + (int64_t)blockIdx.y * args.b_variant_stride;
  const int n0 = blockIdx.y * args.n_variant_stride;

  if (warp >= 5) {
    // ================= producers: im2col gather of the A tile =================
    const int pt = threadIdx.x - 160;          // 0..127
    const int c = pt & 7, rg = pt >> 3;
    constexpr int ES = SRC == SRC_U8 ? 1 : 2;  // source element bytes
    constexpr uint32_t LAG = 2;                // async-copy k-blocks in flight behind the newest one
    uint32_t it = 0;                           // k-block iteration counter (across tiles)
    // loop-invariant pieces of the addressing: swizzled destination of my 8 rows, element->byte factors
    uint32_t dst_off[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) dst_off[p] = swz_off<128>(p * 16 + rg, c);
    const int row_bytes = args.W * args.C * ES;           // one source map row
    const int c_shift = 31 - __clz(args.C);               // C is a power of two (32 / 64 / 512)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      // tap-(0,0) source pointer of each of my 8 rows; one div/mod pair per tile, the rest incremental
      const uint8_t* rptr[8];
      int py0[8], px0[8];
      uint32_t valid = 0;
      {
        const int m0 = tile * 128 + rg;
        int b = m0 / args.P, rem = m0 - b * args.P;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
          const int m = m0 + p * 16;
          const int oy = rem / args.OW, ox = rem - oy * args.OW;
          py0[p] = oy * args.stride - args.pad;
          px0[p] = ox * args.stride - args.pad;
          if (m < args.M) {
            const int64_t srow = args.row_idx ? args.row_idx[b] : (int64_t)b;
            rptr[p] = reinterpret_cast<const uint8_t*>(args.src) +
                      ((srow * args.H + py0[p]) * (int64_t)args.W + px0[p]) * (args.C * ES);
            valid |= 1u << p;
          } else {
            rptr[p] = reinterpret_cast<const uint8_t*>(args.src);
          }
          rem += 16;
          while (rem >= args.P) { rem -= args.P; ++b; }
        }
      }
      int ty = 0, kbx = 0;
      for (int kb = 0; kb < args.KB; ++kb, ++it) {
        const uint32_t stage = it % S, phase = (it / S) & 1;
        mbar_wait(empty + stage, phase ^ 1);
        if (pt == 0) {
          mbar_arrive_expect_tx(full + stage, L::kBBytes);
          tma_bulk_g2s(sB + stage * L::kBBytes, bpack + (int64_t)kb * L::kBBytes, L::kBBytes, full + stage);
        }
        const uint32_t a_addr = smem_u32(sA + stage * L::kABytes);
        if (SRC == SRC_U8) {
          // k-block = 2 kernel rows x (8 px x 4 ch): chunk c -> ky = 2kb + (c>>2), 2 pixels at (c&3)*2
          const int koff = (2 * kb + (c >> 2)) * row_bytes + (c & 3) * 8;
          uint2 v[8];
#pragma unroll
          for (int p = 0; p < 8; ++p) v[p] = (valid >> p) & 1u ? ldg_v2(rptr[p] + koff) : make_uint2(0u, 0u);
#pragma unroll
          for (int p = 0; p < 8; ++p) st_shared_v4(a_addr + dst_off[p], u8x8_to_f16x8(v[p]));
          fence_proxy_async_smem();
          mbar_arrive(full + stage);
        } else {
          const int eoff = kbx * 64 + c * 8;                      // element offset inside the tap row
          const int koff = ty * row_bytes + eoff * ES;
          if (args.pad > 0) {
            const int tx = eoff >> c_shift;                       // pixel offset of this chunk
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const int y = py0[p] + ty, x = px0[p] + tx;
              const bool ok = ((valid >> p) & 1u) && y >= 0 && y < args.H && x >= 0 && x < args.W;
              cp_async_16(a_addr + dst_off[p], ok ? rptr[p] + koff : reinterpret_cast<const uint8_t*>(args.src), ok ? 16u : 0u);
            }
          } else {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              const bool ok = (valid >> p) & 1u;
              cp_async_16(a_addr + dst_off[p], ok ? rptr[p] + koff : rptr[p], ok ? 16u : 0u);
            }
          }
          cp_async_commit();
          if (it >= LAG) {
            cp_async_wait<LAG>();
            fence_proxy_async_smem();
            mbar_arrive(full + ((it - LAG) % S));
          }
          if (++kbx == args.KBX) { kbx = 0; ++ty; }
        }
      }
    }
    if (SRC != SRC_U8) {     // drain the copies still in flight
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (uint32_t j = it >= LAG ? it - LAG : 0; j < it; ++j) mbar_arrive(full + (j % S));
    }
  } else if (warp == 4) {
    // ================= MMA issuer (all lanes stay converged; the mma_* wrappers elect one) =================
    {
      constexpr uint32_t idesc = make_idesc_f16(SRC == SRC_U8 ? 0u : 1u, 128, N, 0, 0);
      constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
      const uint32_t a_lo0 = desc_lo(smem_u32(sA), 0), b_lo0 = desc_lo(smem_u32(sB), 0);
      uint32_t stage = 0, phase = 0, it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
        mbar_wait(tempty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * N;
        for (int kb = 0; kb < args.KB; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (L::kABytes >> 4), b_lo = b_lo0 + stage * (L::kBBytes >> 4);
          if (elect_one()) {
            if (kb == 0) mma_f16_lohi<false>(tmem_d, a_lo, hi, b_lo, hi, idesc);
            else mma_f16_lohi<true>(tmem_d, a_lo, hi, b_lo, hi, idesc);
            mma_f16_lohi<true>(tmem_d, a_lo + 2, hi, b_lo + 2, hi, idesc);
            mma_f16_lohi<true>(tmem_d, a_lo + 4, hi, b_lo + 4, hi, idesc);
            mma_f16_lohi<true>(tmem_d, a_lo + 6, hi, b_lo + 6, hi, idesc);
            mma_commit(empty + stage);
            if (kb == args.KB - 1) mma_commit(tfull + acc);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: TMEM -> registers -> global =================
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
      const int m = tile * 128 + threadIdx.x;       // warps 0..3 own TMEM lanes 32*warp .. +31
      int64_t ooff = -1;
      if (m < args.M) {
        const int b = m / args.P, rem = m - b * args.P;
        const int oy = rem / args.OW, ox = rem - oy * args.OW;
        ooff = (int64_t)b * args.out_sample_pitch + (int64_t)(oy * args.osy) * args.out_row_pitch +
               (int64_t)(ox * args.osx) * args.out_C + n0;
      }
      // chunk ch (32 columns) goes to ooff + coff(ch): consecutive columns, or — for the merged
      // parity classes of the stride-2 transposed conv — output pixel (2u + ch/2, 2v + ch%2)
      auto coff = [&](int ch) -> int64_t {
        return args.classes ? (int64_t)(ch >> 1) * args.out_row_pitch + (int64_t)(ch & 1) * args.out_C : (int64_t)ch * 32;
      };
      uint4 mk[N <= 64 ? N / 8 : 1];
      if (EPI == EPI_MASK_BF16 && N <= 64 && ooff >= 0) {      // prefetch the ReLU mask before the accumulator is ready
#pragma unroll
        for (int ch = 0; ch < N / 32; ++ch)
#pragma unroll
          for (int q = 0; q < 4; ++q) mk[ch * 4 + q] = ldg_v4(args.mask_src + ooff + coff(ch) + 8 * q);
      }
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < N / 32; ++ch) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * N + ch * 32, r);
        tmem_ld_wait();
        if (ch == N / 32 - 1) {
          tc_fence_before();
          mbar_arrive(tempty + acc);               // accumulator drained: MMA may overwrite it
        }
        if (ooff >= 0) {
          if (EPI == EPI_BIAS_RELU_F32) {
            float* o = reinterpret_cast<float*>(args.out) + ooff + ch * 32;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = *reinterpret_cast<const float4*>(sBias + ch * 32 + 4 * q);
              float4 v;
              v.x = fmaxf(fmaf(__uint_as_float(r[4 * q + 0]), args.scale, bb.x), 0.f);
              v.y = fmaxf(fmaf(__uint_as_float(r[4 * q + 1]), args.scale, bb.y), 0.f);
              v.z = fmaxf(fmaf(__uint_as_float(r[4 * q + 2]), args.scale, bb.z), 0.f);
              v.w = fmaxf(fmaf(__uint_as_float(r[4 * q + 3]), args.scale, bb.w), 0.f);
              reinterpret_cast<float4*>(o)[q] = v;
            }
          } else if (EPI == EPI_BIAS_RELU_BF16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + ooff + ch * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j)
                f[j] = fmaxf(fmaf(__uint_as_float(r[8 * q + j]), args.scale, sBias[ch * 32 + 8 * q + j]), 0.f);
              uint4 pk = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
              reinterpret_cast<uint4*>(o)[q] = pk;
            }
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + ooff + coff(ch);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 mv = N <= 64 ? mk[ch * 4 + q] : ldg_v4(args.mask_src + ooff + coff(ch) + 8 * q);
              const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
              uint32_t pk[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // post-ReLU activations are >= 0: "> 0" <=> any non-sign bit set
                const float lo = (mw[j] & 0x00007fffu) ? __uint_as_float(r[8 * q + 2 * j]) : 0.f;
                const float hi = (mw[j] & 0x7fff0000u) ? __uint_as_float(r[8 * q + 2 * j + 1]) : 0.f;
                pk[j] = pack_bf16x2(lo, hi);
              }
              reinterpret_cast<uint4*>(o)[q] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace drl

namespace drl {

// ---- family 2: weight gradients ------------------------------------------------------------------------
struct WgradArgs {
  int M;                    // reduction length: nb * P rows
  // P gather (same geometry vocabulary as RowGatherArgs; no padding in any forward conv)
  const void* src; const int64_t* row_idx;
  int P, OW, H, W, C, stride, KBX;
  int n_atoms_valid;        // Kp / 64 (atoms of the patch that exist)
  // dY
  const __nv_bfloat16* dy;  // [M, dy_ld]
  int dy_ld;
  // work split: blockIdx.x = m-split, blockIdx.y = (M-block group, N tile)
  int mb_per_cta;           // 128-row accumulator blocks handled by one CTA
  int n_tiles;              // N tiles (blockIdx.y % n_tiles)
  // output: fp32 gradient in OIHW layout [Cout][Cin][KH][KW]; kp = (ky, kx, ci)
  float* gw;
  int Cin, KH, KW;
  float scale;              // applied to the accumulator (conv1: the fused 1/255)
};

template <int NT>
struct WgradCfg {
  static constexpr int kRows = 32;                                // reduction rows per smem stage
  static constexpr int kStages = 4;
  static constexpr int kDyRowB = NT >= 64 ? 128 : 64;             // bytes per dY tile row (per atom)
  static constexpr int kDyAtoms = NT >= 64 ? NT / 64 : 1;
  static constexpr int kDyAtomBytes = kRows * kDyRowB;
  static constexpr int kDyBytes = kDyAtoms * kDyAtomBytes;
  static constexpr int kAtomBytes = kRows * 128;                  // one patch atom: 32 rows x 64 kp
};

template <int NT, int SRC, int MB>     // MB: accumulator blocks (128 kp each) per CTA
__global__ void __launch_bounds__(288, (MB == 1 ? 2 : 1)) wgrad_gemm_kernel(const WgradArgs args) {
  using namespace umma;
  using Cf = WgradCfg<NT>;
  constexpr int S = Cf::kStages;
  constexpr int ROWS = Cf::kRows;
  constexpr int NA = 2 * MB;                                    // patch atoms per stage
  constexpr int P_BYTES = NA * Cf::kAtomBytes;
  constexpr int STAGE = P_BYTES + Cf::kDyBytes;
  constexpr int ACC_COLS = MB * NT;
  constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128 : ACC_COLS <= 256 ? 256 : 512;
  static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
  static_assert(STAGE % 1024 == 0, "stage alignment");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * STAGE);
  uint64_t* full = bars;          // [S]
  uint64_t* empty = bars + S;     // [S]
  uint64_t* done = bars + 2 * S;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // this CTA's slice of the reduction, in stages of ROWS rows
  const int total_stages = (args.M + ROWS - 1) / ROWS;
  const int per = (total_stages + gridDim.x - 1) / gridDim.x;
  const int st_begin = blockIdx.x * per;
  const int st_end = min(total_stages, st_begin + per);
  const int n_tile = blockIdx.y % args.n_tiles, mb_group = blockIdx.y / args.n_tiles;
  const int atom0 = mb_group * NA;
  const int n0 = n_tile * NT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 128); mbar_init(empty + s, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (st_begin < st_end) {
    if (warp >= 5) {
      // ================= producers =================
      const int pt = threadIdx.x - 160;
      const int c = pt & 7, slot = pt >> 3;      // rows slot*2, slot*2+1 of the stage
      constexpr int ES = SRC == SRC_U8 ? 1 : 2;
      constexpr uint32_t LAG = 2;
      // loop invariants: per-atom source offset / validity, swizzled destinations of my two rows
      int aoff[NA];
      uint32_t avalid = 0;
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const int atom = atom0 + a;
        if (atom < args.n_atoms_valid) avalid |= 1u << a;
        if (SRC == SRC_U8) {
          aoff[a] = (2 * atom + (c >> 2)) * args.W * args.C + (c & 3) * 8;
        } else {
          const int ty = atom / args.KBX, kbx = atom - ty * args.KBX;
          aoff[a] = (ty * args.W * args.C + kbx * 64 + c * 8) * ES;
        }
      }
      const uint32_t sw128[2] = {swz_off<128>(slot * 2, c), swz_off<128>(slot * 2 + 1, c)};
      const uint32_t sw64[2] = {swz_off<64>(slot * 2, c & 3), swz_off<64>(slot * 2 + 1, c & 3)};
      // (sample, position) of my first row at the first stage; advanced incrementally afterwards
      int b_cur, rem_cur;
      {
        const int m0 = st_begin * ROWS + slot * 2;
        b_cur = m0 / args.P;
        rem_cur = m0 - b_cur * args.P;
      }
      const int64_t pix_bytes = (int64_t)args.C * ES;
      uint32_t it = 0;
      for (int st = st_begin; st < st_end; ++st, ++it) {
        const uint32_t stage = it & (S - 1), phase = (it / S) & 1;
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t* base_s = smem + stage * STAGE;
        const uint32_t p_addr = smem_u32(base_s), d_addr = smem_u32(base_s + P_BYTES);
        int b = b_cur, rem = rem_cur;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int m = st * ROWS + slot * 2 + i;
          const bool live = m < args.M;
          const uint8_t* rptr = reinterpret_cast<const uint8_t*>(args.src);
          if (live) {
            const int oy = rem / args.OW, ox = rem - oy * args.OW;
            const int64_t srow = args.row_idx ? args.row_idx[b] : (int64_t)b;
            rptr += ((srow * args.H + oy * args.stride) * (int64_t)args.W + ox * args.stride) * pix_bytes;
          }
          // dY tile row (bf16, always an async copy)
          const __nv_bfloat16* dyp = args.dy + (live ? (int64_t)m * args.dy_ld + n0 + c * 8 : 0);
          if (NT >= 64) {
#pragma unroll
            for (int a = 0; a < Cf::kDyAtoms; ++a)
              cp_async_16(d_addr + a * Cf::kDyAtomBytes + sw128[i], dyp + (live ? a * 64 : 0), live ? 16u : 0u);
          } else if (c < 4) {
            cp_async_16(d_addr + sw64[i], dyp, live ? 16u : 0u);
          }
          // patch row: atoms atom0 .. atom0+NA-1
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            const bool ok = live && ((avalid >> a) & 1u);
            const uint32_t dst = p_addr + a * Cf::kAtomBytes + sw128[i];
            if (SRC == SRC_U8) {
              uint4 v = make_uint4(0u, 0u, 0u, 0u);
              if (ok) {
                const uint2 w = ldg_v2(rptr + aoff[a]);
                v.x = pack_bf16x2((float)(w.x & 0xffu), (float)((w.x >> 8) & 0xffu));
                v.y = pack_bf16x2((float)((w.x >> 16) & 0xffu), (float)(w.x >> 24));
                v.z = pack_bf16x2((float)(w.y & 0xffu), (float)((w.y >> 8) & 0xffu));
                v.w = pack_bf16x2((float)((w.y >> 16) & 0xffu), (float)(w.y >> 24));
              }
              st_shared_v4(dst, v);
            } else {
              cp_async_16(dst, ok ? rptr + aoff[a] : rptr, ok ? 16u : 0u);
            }
          }
          if (++rem == args.P) { rem = 0; ++b; }
        }
        cp_async_commit();
        if (it >= LAG) {
          cp_async_wait<LAG>();
          fence_proxy_async_smem();
          mbar_arrive(full + ((it - LAG) & (S - 1)));
        }
        rem_cur += ROWS;
        while (rem_cur >= args.P) { rem_cur -= args.P; ++b_cur; }
      }
      cp_async_wait<0>();
      fence_proxy_async_smem();
      for (uint32_t j = it >= LAG ? it - LAG : 0; j < it; ++j) mbar_arrive(full + (j & (S - 1)));
    } else if (warp == 4) {
      // ================= MMA issuer (all lanes stay converged; the mma_* wrappers elect one) =================
      {
        constexpr uint32_t idesc = make_idesc_f16(1u, 128, NT, 1, 1);       // bf16, both operands MN-major
        constexpr uint32_t a_hi = desc_hi(1024, LAYOUT_SW128);
        constexpr uint32_t b_hi = NT >= 64 ? desc_hi(1024, LAYOUT_SW128) : desc_hi(512, LAYOUT_SW64);
        constexpr uint32_t b_kstep = NT >= 64 ? (16 * 128) >> 4 : (16 * 64) >> 4;
        // A: 2 MN-atoms of 64 kp per accumulator block (LBO = atom stride), 16 reduction rows per MMA
        const uint32_t a_lo0 = desc_lo(smem_u32(smem), Cf::kAtomBytes);
        const uint32_t b_lo0 = desc_lo(smem_u32(smem) + P_BYTES, NT >= 64 ? Cf::kDyAtomBytes : 0);
        uint32_t it = 0;
        for (int st = st_begin; st < st_end; ++st, ++it) {
          const uint32_t stage = it & (S - 1), phase = (it / S) & 1;
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (STAGE >> 4), b_lo = b_lo0 + stage * (STAGE >> 4);
          if (elect_one()) {
            if (it == 0) {
#pragma unroll
              for (int mb = 0; mb < MB; ++mb) {
                mma_f16_lohi<false>(tmem_base + mb * NT, a_lo + mb * ((2 * Cf::kAtomBytes) >> 4), a_hi, b_lo, b_hi, idesc);
#pragma unroll
                for (int k = 1; k < ROWS / 16; ++k)
                  mma_f16_lohi<true>(tmem_base + mb * NT, a_lo + mb * ((2 * Cf::kAtomBytes) >> 4) + k * ((16 * 128) >> 4), a_hi,
                                     b_lo + k * b_kstep, b_hi, idesc);
              }
            } else {
#pragma unroll
              for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int k = 0; k < ROWS / 16; ++k)
                  mma_f16_lohi<true>(tmem_base + mb * NT, a_lo + mb * ((2 * Cf::kAtomBytes) >> 4) + k * ((16 * 128) >> 4), a_hi,
                                     b_lo + k * b_kstep, b_hi, idesc);
            }
            mma_commit(empty + stage);
          }
          __syncwarp();
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
      }
      __syncwarp();
    } else {
      // ================= epilogue: TMEM -> red.global.add into the OIHW gradient =================
      mbar_wait(done, 0);
      tc_fence_after();
      const int khw = args.KH * args.KW;
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const int kp = (atom0 + 2 * mb) * 64 + threadIdx.x;       // this thread's accumulator row
        const bool ok = kp < args.n_atoms_valid * 64;
        // kp = (ky, kx, ci)
        const int kc = args.KW * args.Cin;
        const int ky = kp / kc, r2 = kp - ky * kc;
        const int kx = r2 / args.Cin, ci = r2 - kx * args.Cin;
        float* g0 = args.gw + ((int64_t)ci * args.KH + ky) * args.KW + kx;
#pragma unroll
        for (int ch = 0; ch < NT / 32; ++ch) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * NT + ch * 32, r);
          tmem_ld_wait();
          if (ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int co = n0 + ch * 32 + j;
              atomicAdd(g0 + (int64_t)co * args.Cin * khw, __uint_as_float(r[j]) * args.scale);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace drl
