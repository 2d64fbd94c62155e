// conv1 of NatureCNN (8x8 stride 4 over uint8 84x84x4 frames -> 20x20x32) as tcgen05 kernels that read
// the uint8 frames DIRECTLY: the frame rows a tile needs are contiguous in memory (NHWC rows of
// 336 B), so they are staged into shared memory by 1-D TMA bulk copies (gathered through the
// minibatch row index), converted u8 -> fp16/bf16 on the way into the swizzled operand tiles by the
// producer warps, and the x*(1/255) of scale_observations.py:35 is applied to the accumulator.
//
//   conv1_fwd_kernel   D[128 positions x 32] = im2col(frame)[128 x 256] * W1[32 x 256]^T
//   conv1_wgrad_kernel dW1^T[256 x 32] += im2col(frame)[m x 256]^T * dY1[m x 32]
//
// A tile is 128 consecutive rows of the flattened (sample, oy, ox) grid; it touches at most two
// samples ("segments"), each needing input rows [4*oy_a, 4*oy_b + 8) = at most 36 rows = 12 096 B.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace drl {

struct Conv1Args {
  const uint8_t* obs;        // [rows, 84, 84, 4]
  const int64_t* row_idx;    // [nb] or nullptr
  int nb;
  const void* bpack;         // fwd: fp16 [4][32 x 64] swizzled tile images (16 KB)
  const float* bias;         // fwd: [32] EFFECTIVE bias = b - 1024 * scale * sum_k fp16(W[n][k]) (operands hold 1024 + pixel)
  float scale;               // float32(1/255)
  __nv_bfloat16* out;        // fwd: act1 [nb*400, 32]
  uint32_t* mask_bits;       // fwd: ReLU mask word of position m (umma.cuh relu_bit_pos): what conv2's data gradient needs
  const __nv_bfloat16* dy;   // wgrad: dY1 [nb*400, 32]
  float* gw;                 // wgrad: fp32 [32][4][8][8]
};

constexpr int kC1SegBytes = 12288;
constexpr int kC1FrameBytes = 84 * 84 * 4;
constexpr int kC1RowBytes = 84 * 4;

// p / 400 for 0 <= p < 2^24 without an integer division (float reciprocal + one-step correction)
__device__ __forceinline__ int c1_div400(int p, int& rem) {
  int q = (int)((float)p * 0.0025f);
  rem = p - q * 400;
  if (rem < 0) { rem += 400; --q; }
  if (rem >= 400) { rem -= 400; ++q; }
  return q;
}
// x / 20 for 0 <= x < 400 (exact: 3277 = ceil(2^16 / 20))
__device__ __forceinline__ int c1_div20(int x) { return (x * 3277) >> 16; }

struct C1Seg { int rows_begin; int oy_a; uint32_t bytes; int64_t src_off; bool valid; };

// Segment s (0/1) of the tile starting at flattened position p0.
__device__ __forceinline__ void c1_segment(int p0, int s, int nb, const int64_t* row_idx, C1Seg& g) {
  int r0;
  const int b0 = c1_div400(p0, r0);
  int b, ra, rb;
  if (s == 0) { b = b0; ra = r0; rb = min(399, r0 + 127); }
  else { b = b0 + 1; ra = 0; rb = r0 + 127 - 400; }
  g.valid = b < nb && rb >= ra;
  if (!g.valid) { g.bytes = 0; g.oy_a = 0; g.src_off = 0; g.rows_begin = 0; return; }
  const int oy_a = c1_div20(ra), oy_b = c1_div20(rb);
  g.oy_a = oy_a;
  g.bytes = (uint32_t)(((oy_b - oy_a) * 4 + 8) * kC1RowBytes);
  const int64_t srow = row_idx ? row_idx[b] : (int64_t)b;
  g.src_off = srow * kC1FrameBytes + (int64_t)oy_a * 4 * kC1RowBytes;
}

// smem byte offset (inside the raw staging buffer pair of one tile) of the tap-(0,0) pixel of tile
// row r, or -1 if the row lies past the end of the batch.
__device__ __forceinline__ int c1_row_base(int p0, int r, int nb) {
  int rem0, rem;
  const int b0 = c1_div400(p0, rem0);
  const int b = c1_div400(p0 + r, rem);
  if (b >= nb) return -1;
  const int oy = c1_div20(rem), ox = rem - oy * 20;
  const int oy_a = b == b0 ? c1_div20(rem0) : 0;
  return (b - b0) * kC1SegBytes + (oy - oy_a) * 4 * kC1RowBytes + ox * 16;
}

__device__ __forceinline__ uint2 lds_v2(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}

__device__ __forceinline__ uint4 lds_v4(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

// 8 uint8 -> 8 fp16 holding 1024 + b (bytes become the mantissa of 0x6400 = 1024.0): one PRMT per pair.
// The constant 1024 * sum_k W[n][k] this adds to every accumulator is folded into the bias
// (last block of pack_all_kernel), so the forward conversion costs 4 instructions per 8 pixels.
__device__ __forceinline__ uint4 u8x8_to_f16x8_biased(uint2 w) {
  const uint32_t magic = 0x64006400u;
  return make_uint4(__byte_perm(w.x, magic, 0x7170), __byte_perm(w.x, magic, 0x7372),
                    __byte_perm(w.y, magic, 0x7170), __byte_perm(w.y, magic, 0x7372));
}

// 8 uint8 -> 8 bf16 (exact: 0..255 fit the 8-bit significand): fp32 magic number then truncation.
__device__ __forceinline__ uint4 u8x8_to_bf16x8(uint2 w) {
  const uint32_t magic = 0x4B000000u;                 // 2^23
  float f[8];
  f[0] = __uint_as_float(__byte_perm(w.x, magic, 0x7440)) - 8388608.f;   // {b0,0,0,0x4B}
  f[1] = __uint_as_float(__byte_perm(w.x, magic, 0x7441)) - 8388608.f;
  f[2] = __uint_as_float(__byte_perm(w.x, magic, 0x7442)) - 8388608.f;
  f[3] = __uint_as_float(__byte_perm(w.x, magic, 0x7443)) - 8388608.f;
  f[4] = __uint_as_float(__byte_perm(w.y, magic, 0x7440)) - 8388608.f;
  f[5] = __uint_as_float(__byte_perm(w.y, magic, 0x7441)) - 8388608.f;
  f[6] = __uint_as_float(__byte_perm(w.y, magic, 0x7442)) - 8388608.f;
  f[7] = __uint_as_float(__byte_perm(w.y, magic, 0x7443)) - 8388608.f;
  uint4 o;   // bf16x2 = high halves of the two floats
  o.x = __byte_perm(__float_as_uint(f[0]), __float_as_uint(f[1]), 0x7632);
  o.y = __byte_perm(__float_as_uint(f[2]), __float_as_uint(f[3]), 0x7632);
  o.z = __byte_perm(__float_as_uint(f[4]), __float_as_uint(f[5]), 0x7632);
  o.w = __byte_perm(__float_as_uint(f[6]), __float_as_uint(f[7]), 0x7632);
  return o;
}

// conv1 epilogue of one output position: relu(acc * (1/255) + bias) -> 32 bf16 (one 64-byte row) + its ReLU mask word
__device__ __forceinline__ void c1_store_row(const uint32_t (&r)[32], const float (&bias)[32], float scale,
                                             __nv_bfloat16* out_row, uint32_t* bits_dst) {
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    pk[i] = umma::pack_bf16x2_relu(fmaf(__uint_as_float(r[2 * i]), scale, bias[2 * i]),
                                   fmaf(__uint_as_float(r[2 * i + 1]), scale, bias[2 * i + 1]));
#pragma unroll
  for (int q = 0; q < 2; ++q)
    umma::stg_v8(out_row + 16 * q, pk[8 * q], pk[8 * q + 1], pk[8 * q + 2], pk[8 * q + 3], pk[8 * q + 4], pk[8 * q + 5],
                 pk[8 * q + 6], pk[8 * q + 7]);
  *bits_dst = umma::relu_bits_from_pairs(pk);
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
constexpr int kC1ProdWarps = 8;
constexpr int kC1Threads = 32 * (5 + kC1ProdWarps + 1);   // 4 epilogue + 1 MMA + 8 converters + 1 TMA warp
constexpr int kC1NR = 3;                                    // raw staging buffers (tiles in flight)
constexpr int kC1FwdSmem = 16384 /*B*/ + 2 * 65536 /*A: two whole tiles*/ + kC1NR * 2 * kC1SegBytes /*raw*/ + 256 + 1024;

__global__ void __launch_bounds__(kC1Threads, 1) conv1_fwd_kernel(const Conv1Args args) {
  using namespace umma;
  constexpr int NP = 32 * kC1ProdWarps;          // producer threads
  constexpr int RPT = 128 / (NP / 8);            // tile rows per producer thread
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sA = smem + 16384;
  uint8_t* sRaw = smem + 16384 + 2 * 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + kC1NR * 2 * kC1SegBytes);
  uint64_t* full = bars;            // [2] one per A buffer (a whole 128 x 256 tile = 4 k-block images)
  uint64_t* empty = bars + 4;       // [2]
  uint64_t* tfull = bars + 8;       // [2]
  uint64_t* tempty = bars + 10;     // [2]
  uint64_t* raw_full = bars + 12;   // [kC1NR]
  uint64_t* raw_empty = bars + 16;  // [kC1NR]
  uint64_t* b_full = bars + 20;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = args.nb * 400;
  const int num_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(full + s, NP); mbar_init(empty + s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 128); }
    for (int a = 0; a < kC1NR; ++a) { mbar_init(raw_full + a, 1); mbar_init(raw_empty + a, NP); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5 + kC1ProdWarps) {
    // ================= TMA warp: stage the raw frame rows of up to kC1NR tiles ahead ====================
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, 16384);
      tma_bulk_g2s(sB, args.bpack, 16384, b_full);
      uint32_t j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
        const uint32_t buf = j % kC1NR;
        mbar_wait(raw_empty + buf, ((j / kC1NR) & 1) ^ 1);
        C1Seg s0, s1;
        c1_segment(tile * 128, 0, args.nb, args.row_idx, s0);
        c1_segment(tile * 128, 1, args.nb, args.row_idx, s1);
        mbar_arrive_expect_tx(raw_full + buf, s0.bytes + s1.bytes);
        uint8_t* dst = sRaw + buf * 2 * kC1SegBytes;
        if (s0.valid) tma_bulk_g2s(dst, args.obs + s0.src_off, s0.bytes, raw_full + buf);
        if (s1.valid) tma_bulk_g2s(dst + kC1SegBytes, args.obs + s1.src_off, s1.bytes, raw_full + buf);
      }
    }
    __syncwarp();
  } else if (warp >= 5) {
    // ================= converters: raw u8 rows -> fp16 im2col rows of the 4 swizzled A tiles =============
    const int pt = threadIdx.x - 160;
    const int c = pt & 7, rg = pt >> 3;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const uint32_t buf = j % kC1NR;
      int rb[RPT];
#pragma unroll
      for (int q = 0; q < RPT; ++q) rb[q] = c1_row_base(tile * 128, q * (NP / 8) + rg, args.nb);
      mbar_wait(raw_full + buf, (j / kC1NR) & 1);
      const uint32_t raw_addr = smem_u32(sRaw + buf * 2 * kC1SegBytes);
      const uint32_t ab = j & 1;
      mbar_wait(empty + ab, ((j >> 1) & 1) ^ 1);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint32_t a_addr = smem_u32(sA + ab * 65536 + kb * 16384);
        const uint32_t koff = (uint32_t)((2 * kb + (c >> 2)) * kC1RowBytes + (c & 3) * 8);
        uint2 v[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q) v[q] = rb[q] >= 0 ? lds_v2(raw_addr + rb[q] + koff) : make_uint2(0u, 0u);
#pragma unroll
        for (int q = 0; q < RPT; ++q) st_shared_v4(a_addr + swz_off<128>(q * (NP / 8) + rg, c), u8x8_to_f16x8_biased(v[q]));
      }
      fence_proxy_async_smem();
      mbar_arrive(full + ab);
      mbar_arrive(raw_empty + buf);
    }
  } else if (warp == 4) {
    {
      constexpr uint32_t idesc = make_idesc_f16(0u, 128, 32, 0, 0);     // fp16 operands
      constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
      const uint32_t a_lo0 = desc_lo(smem_u32(sA), 0), b_lo0 = desc_lo(smem_u32(sB), 0);
      mbar_wait(b_full, 0);
      uint32_t j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
        const uint32_t acc = j & 1, ab = j & 1;
        mbar_wait(tempty + acc, ((j >> 1) & 1) ^ 1);
        mbar_wait(full + ab, (j >> 1) & 1);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + ab * (65536 >> 4);
        if (elect_one()) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kb == 0 && k == 0)
                mma_f16_lohi<false>(tmem_base + acc * 32, a_lo, hi, b_lo0, hi, idesc);
              else
                mma_f16_lohi<true>(tmem_base + acc * 32, a_lo + kb * (16384 >> 4) + k * 2, hi, b_lo0 + kb * (4096 >> 4) + k * 2, hi, idesc);
            }
          }
          mma_commit(empty + ab);
          mma_commit(tfull + acc);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue: acc * (1/255) + bias, ReLU, bf16, one 64-byte row per thread ===========
    float bias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias[i] = __ldg(args.bias + i);
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const uint32_t acc = j & 1;
      const int m = tile * 128 + threadIdx.x;
      mbar_wait(tfull + acc, (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tempty + acc);
      if (m < M) c1_store_row(r, bias, args.scale, args.out + (int64_t)m * 32, args.mask_bits + m);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward, A operand in TENSOR MEMORY.  The im2col row of an output position is 8 runs of 32 contiguous
// bytes of the staged frame rows, so one converter thread owns one tile row (= one TMEM lane): it reads its
// runs with LDS.128, turns them into 1024-biased fp16 pairs and writes them with tcgen05.st into the
// [128 lanes x 128 columns] A buffer; the MMA then takes A from TMEM and only the 16 KB of weights from
// shared memory.  Per tile the shared-memory traffic drops from ~176 KB (raw read + swizzled fp16 write +
// operand read) to ~72 KB, and the generic->async proxy fence disappears from the converter path.
// TMEM columns: [0,128) four accumulators, then three 128-column A buffers (all 512 columns, one CTA per SM).
// ---------------------------------------------------------------------------------------------------------
constexpr int kC1tNR = 4;                                   // raw staging buffers (tiles in flight)
constexpr int kC1tNA = 3;                                   // A buffers in tensor memory
constexpr int kC1tNACC = 4;                                 // accumulators
constexpr int kC1tEpiWG = 2;                                // epilogue warpgroups (tiles round-robin)
constexpr int kC1tThreads = 32 * (4 * kC1tEpiWG + 1 + kC1ProdWarps + 1);
constexpr int kC1FwdTsSmem = 16384 /*B*/ + kC1tNR * 2 * kC1SegBytes /*raw*/ + 256 + 1024;

__global__ void __launch_bounds__(kC1tThreads, 1) conv1_fwd_ts_kernel(const Conv1Args args) {
  using namespace umma;
  constexpr int NP = 32 * kC1ProdWarps;          // converter threads
  constexpr int W_MMA = 4 * kC1tEpiWG, W_CONV = W_MMA + 1, W_TMA = W_CONV + kC1ProdWarps;
  constexpr uint32_t ACC_COL = 0, A_COL = kC1tNACC * 32;   // tensor-memory columns: accumulators, then the A buffers
  static_assert(A_COL + kC1tNA * 128 <= 512, "tensor memory budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;
  uint8_t* sRaw = smem + 16384;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + kC1tNR * 2 * kC1SegBytes);
  uint64_t* full = bars;                         // [NA] A buffer written
  uint64_t* empty = full + kC1tNA;               // [NA] A buffer consumed by the MMAs
  uint64_t* tfull = empty + kC1tNA;              // [NACC]
  uint64_t* tempty = tfull + kC1tNACC;           // [NACC]
  uint64_t* raw_full = tempty + kC1tNACC;        // [NR]
  uint64_t* raw_empty = raw_full + kC1tNR;       // [NR]
  uint64_t* b_full = raw_empty + kC1tNR;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);
  static_assert((2 * kC1tNA + 2 * kC1tNACC + 2 * kC1tNR + 2) * 8 <= 256, "barrier block");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = args.nb * 400;
  const int num_tiles = (M + 127) / 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kC1tNA; ++s) { mbar_init(full + s, NP); mbar_init(empty + s, 1); }
    for (int a = 0; a < kC1tNACC; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 128); }
    for (int a = 0; a < kC1tNR; ++a) { mbar_init(raw_full + a, 1); mbar_init(raw_empty + a, NP); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0) {
      mbar_arrive_expect_tx(b_full, 16384);
      tma_bulk_g2s(sB, args.bpack, 16384, b_full);
      uint32_t j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
        const uint32_t buf = j % kC1tNR;
        mbar_wait(raw_empty + buf, ((j / kC1tNR) & 1) ^ 1);
        C1Seg s0, s1;
        c1_segment(tile * 128, 0, args.nb, args.row_idx, s0);
        c1_segment(tile * 128, 1, args.nb, args.row_idx, s1);
        mbar_arrive_expect_tx(raw_full + buf, s0.bytes + s1.bytes);
        uint8_t* dst = sRaw + buf * 2 * kC1SegBytes;
        if (s0.valid) tma_bulk_g2s(dst, args.obs + s0.src_off, s0.bytes, raw_full + buf);
        if (s1.valid) tma_bulk_g2s(dst + kC1SegBytes, args.obs + s1.src_off, s1.bytes, raw_full + buf);
      }
    }
    __syncwarp();
  } else if (warp >= W_CONV) {
    // ================= converters: one tile row (TMEM lane) per thread, half of K per warp ==============
    const int quarter = warp & 3;                 // the TMEM lanes this warp may access
    const int khalf = (warp - W_CONV) >> 2;       // ky in [4*khalf, 4*khalf + 4)
    const int row = quarter * 32 + lane;
    const uint32_t t_lane = (uint32_t)(quarter * 32) << 16;
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      const uint32_t buf = j % kC1tNR, ab = j % kC1tNA;
      // rows past the end of the batch read row 0 of the staging buffer: their outputs are never stored
      const int rb = max(c1_row_base(tile * 128, row, args.nb), 0);
      const uint32_t src = smem_u32(sRaw + buf * 2 * kC1SegBytes) + (uint32_t)rb + (uint32_t)(khalf * 4 * kC1RowBytes);
      mbar_wait(raw_full + buf, (j / kC1tNR) & 1);
      uint4 v[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        v[2 * q] = lds_v4(src + (uint32_t)(q * kC1RowBytes));
        v[2 * q + 1] = lds_v4(src + (uint32_t)(q * kC1RowBytes) + 16);
      }
      mbar_wait(empty + ab, ((j / kC1tNA) & 1) ^ 1);
      tc_fence_after();
      const uint32_t t_a = tmem_base + t_lane + A_COL + ab * 128 + khalf * 64;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t r[32];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 w = v[4 * h2 + q];
          const uint4 a = u8x8_to_f16x8_biased(make_uint2(w.x, w.y)), b = u8x8_to_f16x8_biased(make_uint2(w.z, w.w));
          r[8 * q] = a.x; r[8 * q + 1] = a.y; r[8 * q + 2] = a.z; r[8 * q + 3] = a.w;
          r[8 * q + 4] = b.x; r[8 * q + 5] = b.y; r[8 * q + 6] = b.z; r[8 * q + 7] = b.w;
        }
        tmem_st_32x32b_x32(t_a + h2 * 32, r);
      }
      mbar_arrive(raw_empty + buf);                // the raw rows are in registers / tensor memory now
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(full + ab);
    }
  } else if (warp == W_MMA) {
    {
      constexpr uint32_t idesc = make_idesc_f16(0u, 128, 32, 0, 0);     // fp16 operands
      constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
      const uint32_t b_lo0 = desc_lo(smem_u32(sB), 0);
      mbar_wait(b_full, 0);
      uint32_t j = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
        const uint32_t acc = j % kC1tNACC, ab = j % kC1tNA;
        mbar_wait(tempty + acc, ((j / kC1tNACC) & 1) ^ 1);
        mbar_wait(full + ab, (j / kC1tNA) & 1);
        tc_fence_after();
        const uint32_t t_a = tmem_base + A_COL + ab * 128, t_d = tmem_base + ACC_COL + acc * 32;
        if (elect_one()) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint32_t b_lo = b_lo0 + (i >> 2) * (4096 >> 4) + (i & 3) * 2;
            if (i == 0) mma_f16_ts<false>(t_d, t_a, b_lo, hi, idesc);
            else mma_f16_ts<true>(t_d, t_a + i * 8, b_lo, hi, idesc);
          }
          mma_commit(empty + ab);
          mma_commit(tfull + acc);
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warpgroups: tiles round-robin =========================================
    const int wg = warp >> 2, trow = threadIdx.x & 127;
    float bias[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bias[i] = __ldg(args.bias + i);
    uint32_t j = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++j) {
      if ((int)(j % kC1tEpiWG) != wg) continue;
      const uint32_t acc = j % kC1tNACC;
      const int m = tile * 128 + trow;
      mbar_wait(tfull + acc, (j / kC1tNACC) & 1);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + ACC_COL + acc * 32, r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tempty + acc);
      if (m < M) c1_store_row(r, bias, args.scale, args.out + (int64_t)m * 32, args.mask_bits + m);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight gradient:  dW1^T[kp=256 x co=32] += P[m x 256]^T * dY1[m x 32]   (both operands MN-major)
// Per tile of 128 positions: frames + the 8 KB of dY1 rows arrive by TMA into a raw staging buffer;
// producers build 4 sub-stages of 32 rows: 4 patch atoms (32 x 64 kp, bf16, SW128) + the dY tile
// (32 x 32, SW64).  The two 128x32 accumulators live in TMEM for the whole kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int kC1wStage = 4 * 4096 + 2048;                 // 18 KB
constexpr int kC1wRaw = 2 * kC1SegBytes + 8192;            // frames (2 segments) + dY rows
constexpr int kC1wNR = 2;                                   // raw buffers of the wgrad kernel
constexpr int kC1WgradSmem = 8 * kC1wStage + kC1wNR * kC1wRaw + 256 + 1024;

__global__ void __launch_bounds__(kC1Threads, 1) conv1_wgrad_kernel(const Conv1Args args) {
  using namespace umma;
  constexpr int NP = 32 * kC1ProdWarps;          // 256 producer threads
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sSt = smem;                                   // 2 tiles x 4 sub-stages
  uint8_t* sRaw = smem + 8 * kC1wStage;                  // kC1wNR raw buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + kC1wNR * kC1wRaw);
  uint64_t* full = bars;            // [2] one per tile buffer
  uint64_t* empty = bars + 4;       // [2]
  uint64_t* raw_full = bars + 8;    // [kC1wNR]
  uint64_t* raw_empty = bars + 12;  // [kC1wNR]
  uint64_t* done = bars + 16;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = args.nb * 400;
  const int num_tiles = (M + 127) / 128;
  // contiguous slice of tiles per CTA
  const int per = (num_tiles + gridDim.x - 1) / gridDim.x;
  const int t_begin = blockIdx.x * per, t_end = min(num_tiles, t_begin + per);

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(full + s, NP); mbar_init(empty + s, 1); }
    for (int a = 0; a < kC1wNR; ++a) { mbar_init(raw_full + a, 1); mbar_init(raw_empty + a, NP); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(tmem_slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (t_begin < t_end) {
    if (warp == 5 + kC1ProdWarps) {
      if (lane == 0) {
        uint32_t j = 0;
        for (int tile = t_begin; tile < t_end; ++tile, ++j) {
          const uint32_t buf = j % kC1wNR;
          mbar_wait(raw_empty + buf, ((j / kC1wNR) & 1) ^ 1);
          C1Seg s0, s1;
          c1_segment(tile * 128, 0, args.nb, args.row_idx, s0);
          c1_segment(tile * 128, 1, args.nb, args.row_idx, s1);
          const uint32_t dy_bytes = (uint32_t)min(128, M - tile * 128) * 64u;
          mbar_arrive_expect_tx(raw_full + buf, s0.bytes + s1.bytes + dy_bytes);
          uint8_t* dst = sRaw + buf * kC1wRaw;
          if (s0.valid) tma_bulk_g2s(dst, args.obs + s0.src_off, s0.bytes, raw_full + buf);
          if (s1.valid) tma_bulk_g2s(dst + kC1SegBytes, args.obs + s1.src_off, s1.bytes, raw_full + buf);
          tma_bulk_g2s(dst + 2 * kC1SegBytes, args.dy + (int64_t)tile * 128 * 32, dy_bytes, raw_full + buf);
        }
      }
      __syncwarp();
    } else if (warp >= 5) {
      const int pt = threadIdx.x - 160;
      const int c = pt & 7, row = pt >> 3;         // one row of each 32-row sub-stage per thread
      uint32_t j = 0;
      for (int tile = t_begin; tile < t_end; ++tile, ++j) {
        const uint32_t buf = j % kC1wNR;
        int rb[4];
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) rb[ss] = c1_row_base(tile * 128, ss * 32 + row, args.nb);
        mbar_wait(raw_full + buf, (j / kC1wNR) & 1);
        const uint32_t raw_addr = smem_u32(sRaw + buf * kC1wRaw);
        const uint32_t tb = j & 1;
        mbar_wait(empty + tb, ((j >> 1) & 1) ^ 1);
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
          const uint32_t st_addr = smem_u32(sSt + (tb * 4 + ss) * kC1wStage);
          uint2 v[4];
#pragma unroll
          for (int a = 0; a < 4; ++a)
            v[a] = rb[ss] >= 0 ? lds_v2(raw_addr + rb[ss] + (uint32_t)((2 * a + (c >> 2)) * kC1RowBytes + (c & 3) * 8))
                               : make_uint2(0u, 0u);
          uint4 dyv = make_uint4(0u, 0u, 0u, 0u);
          if (c < 4 && rb[ss] >= 0) dyv = lds_v4(raw_addr + 2 * kC1SegBytes + (ss * 32 + row) * 64 + c * 16);
#pragma unroll
          for (int a = 0; a < 4; ++a) st_shared_v4(st_addr + a * 4096 + swz_off<128>(row, c), u8x8_to_bf16x8(v[a]));
          if (c < 4) st_shared_v4(st_addr + 4 * 4096 + swz_off<64>(row, c), dyv);
        }
        fence_proxy_async_smem();
        mbar_arrive(full + tb);
        mbar_arrive(raw_empty + buf);
      }
    } else if (warp == 4) {
      {
        constexpr uint32_t idesc = make_idesc_f16(1u, 128, 32, 1, 1);   // bf16, MN-major A and B
        constexpr uint32_t a_hi = desc_hi(1024, LAYOUT_SW128), b_hi = desc_hi(512, LAYOUT_SW64);
        const uint32_t st_lo0 = desc_lo(smem_u32(sSt), 0);
        uint32_t j = 0;
        for (int tile = t_begin; tile < t_end; ++tile, ++j) {
          const uint32_t tb = j & 1;
          mbar_wait(full + tb, (j >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
          for (int ss = 0; ss < 4; ++ss) {
            const uint32_t p_lo = st_lo0 + (tb * 4 + ss) * (kC1wStage >> 4), d_lo = p_lo + ((4 * 4096) >> 4);
#pragma unroll
            for (int mb = 0; mb < 2; ++mb)
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const uint32_t al = (p_lo + mb * ((2 * 4096) >> 4) + k * ((16 * 128) >> 4)) | (((4096u >> 4) & 0x3FFFu) << 16);
                const uint32_t bl = d_lo + k * ((16 * 64) >> 4);
                if (ss == 0 && k == 0) {
                  if (j == 0) mma_f16_lohi<false>(tmem_base + mb * 32, al, a_hi, bl, b_hi, idesc);
                  else mma_f16_lohi<true>(tmem_base + mb * 32, al, a_hi, bl, b_hi, idesc);
                } else {
                  mma_f16_lohi<true>(tmem_base + mb * 32, al, a_hi, bl, b_hi, idesc);
                }
              }
          }
          mma_commit(empty + tb);
          }
          __syncwarp();
        }
        if (elect_one()) mma_commit(done);
        __syncwarp();
      }
      __syncwarp();
    } else {
      mbar_wait(done, 0);
      tc_fence_after();
#pragma unroll
      for (int mb = 0; mb < 2; ++mb) {
        const int kp = mb * 128 + threadIdx.x;            // (ky, kx, ci): ky = kp/32, kx = (kp%32)/4, ci = kp%4
        const int ky = kp >> 5, kx = (kp & 31) >> 2, ci = kp & 3;
        float* g0 = args.gw + (ci * 8 + ky) * 8 + kx;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + mb * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int co = 0; co < 32; ++co) atomicAdd(g0 + co * 256, __uint_as_float(r[co]) * args.scale);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace drl
