// Family 1e — `tma_conv2_kernel`: the descriptor-shifted implicit-GEMM convolution of umma_tma_conv.cuh on CTA PAIRS
// (tcgen05 `cta_group::2`, M = 256 = one sample per CTA).
//
// Why: these small-N layers are bound by the operand fetch from shared memory (128 B/clk per SM): an M = 128 x N MMA
// step reads 4 KB of activations and 32 N bytes of weights, (4096 + 32 N) / 128 cycles.  In a CTA pair every CTA
// stages ITS sample and only HALF of the weight rows (N / 2 output channels of every tap); the pair's MMA reads
// both halves, so each SM fetches 4096 + 16 N bytes per step: N = 64: 48 -> 40 cycles, N = 128: 64 -> 48 cycles.
// The resident weights also take half the shared memory.
//
// Cluster (2,1,1).  Warps 0..4*NEPI-1 epilogue (each CTA drains its own 128 accumulator rows), warp 4*NEPI MMA
// (leader CTA issues for the pair; both CTAs allocate TMEM with cta_group::2), warp 4*NEPI+1 TMA (each CTA loads its
// own sample; the transaction bytes of BOTH land on the leader's `raw_full`).
//   raw_full[s]   leader only: 1 arrival (expect_tx of both boxes) + the loads of the pair
//   raw_empty[s]  both CTAs: tcgen05.commit multicast after the MMAs that read slot s
//   tfull[a]      both CTAs: tcgen05.commit multicast after the last MMA of a tile
//   tempty[a]     leader only: 8 arrivals = the 4 warps of one epilogue warpgroup of each CTA (one remote arrive per warp)
// Pair p processes sample pairs q = p, p + pairs, ...: rank r owns sample 2q + r (the last pair of an odd batch
// re-loads sample nb - 1 in rank 1 and discards it).
#pragma once
#include "umma_tma_conv.cuh"
#include "umma_tma_gemm2.cuh"

namespace drl {

__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3,
                                                uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tmap), "r"(umma::smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int N>
constexpr int tma_conv2_smem(int KB, int raw_buf_bytes, int NR, int slack) {
  return KB * (N / 2) * 128 + NR * raw_buf_bytes + slack + 1280 + 1024;
}

template <int N, int EPI, int NR, int NACC, int NEPI, int KB, int KBX, int SHIFT_RW, int PLANES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (4 * NEPI + 2), 1)
    tma_conv2_kernel(const __grid_constant__ CUtensorMap tmap, const TmaConvArgs args) {
  using namespace umma;
  constexpr int BH_BYTES = (N / 2) * 128;                     // this CTA's half of one weight k-block tile
  constexpr int TMEM_COLS = NACC * N <= 64 ? 64 : NACC * N <= 128 ? 128 : NACC * N <= 256 ? 256 : 512;
  static_assert(NACC * N <= 512, "accumulators exceed tensor memory");
  static_assert((N / 2) % 8 == 0, "weight halves must keep the 8-row swizzle groups");
  constexpr int W_MMA = 4 * NEPI, W_TMA = 4 * NEPI + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                         // KB half weight tiles, resident
  uint8_t* sRaw = smem + KB * BH_BYTES;                       // NR staged samples
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + NR * args.raw_buf_bytes + args.slack_bytes);
  uint64_t* raw_full = bars;                    // [NR]
  uint64_t* raw_empty = bars + NR;              // [NR]
  uint64_t* tfull = bars + 2 * NR;              // [NACC]
  uint64_t* tempty = bars + 2 * NR + NACC;      // [NACC]
  uint64_t* b_full = bars + 2 * NR + 2 * NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NR + 2 * NACC + 1);
  static_assert((2 * NR + 2 * NACC + 2) * 8 <= 256, "barrier block overflows");
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int total_pairs = (args.nb + 1) >> 1;
  const int n_mine = pair < total_pairs ? (total_pairs - pair + n_pairs - 1) / n_pairs : 0;   // sample pairs of this CTA pair
  auto sample_of = [&](int j) { return 2 * (pair + j * n_pairs) + (int)rank; };

  if (EPI != EPI_MASK_BF16 && threadIdx.x < N) sBias[threadIdx.x] = args.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int a = 0; a < NR; ++a) { mbar_init(raw_full + a, 1); mbar_init(raw_empty + a, 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 8); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc2(tmem_slot, TMEM_COLS);
  if (warp == W_TMA && lane == 0) tma_prefetch_desc(&tmap);
  tc_fence_before();
  cluster_sync_all();                  // barriers of both CTAs initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto load_sample = [&](int j) {      // TMA thread: stage sample pair j (this CTA's half) into slot j % NR
    const uint32_t buf = (uint32_t)j % NR;
    int b = sample_of(j);
    if (b >= args.nb) b = args.nb - 1;
    if (rank == 0) mbar_arrive_expect_tx(raw_full + buf, (uint32_t)(2 * args.box_bytes * PLANES));
    for (int pl = 0; pl < PLANES; ++pl)
      tma_load_4d_2sm(sRaw + buf * args.raw_buf_bytes + pl * args.plane_bytes, &tmap, 0, -args.pad, -args.pad + pl, b, raw_full + buf);
  };
  const int n_pre = n_mine < NR ? n_mine : NR;
  if (warp == W_TMA && lane == 0) {
    // this CTA's half of every weight tile (rows rank * N/2 ... of the [N x 64] swizzled image), resident
    mbar_arrive_expect_tx(b_full, (uint32_t)(KB * BH_BYTES));
    for (int kb = 0; kb < KB; ++kb)
      tma_bulk_g2s(sB + kb * BH_BYTES,
                   reinterpret_cast<const uint8_t*>(args.bpack) + (int64_t)kb * (N * 128) + (int64_t)rank * BH_BYTES, BH_BYTES, b_full);
    for (int j = 0; j < n_pre; ++j) load_sample(j);           // the first ring fill needs no empty-wait
  }
  __syncwarp();
  mbar_wait(b_full, 0);
  cluster_sync_all();                  // BOTH halves of the weights have landed before the leader's first MMA
  tc_fence_after();

  if (warp == W_TMA) {
    if (lane == 0) {
      for (int j = n_pre; j < n_mine; ++j) {
        mbar_wait(raw_empty + (uint32_t)j % NR, (((uint32_t)j / NR) & 1) ^ 1);
        load_sample(j);
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(1u, 256, N, 0, 0);
      constexpr uint32_t hi0 = desc_hi(1024, LAYOUT_SW128);
      const uint32_t b_lo0 = desc_lo(smem_u32(sB), 0);
      // two sample pairs are issued interleaved (k-block by k-block) into two accumulators, as in the single-CTA kernel
      for (int j = 0; j < n_mine; j += 2) {
        const int np = min(2, n_mine - j);
        uint32_t raw_addr[2], tmem_d[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const uint32_t jj = (uint32_t)(j + (t < np ? t : 0)), buf = jj % NR, acc = jj % NACC;
          mbar_wait(raw_full + buf, (jj / NR) & 1);
          mbar_wait(tempty + acc, ((jj / NACC) & 1) ^ 1);
          raw_addr[t] = smem_u32(sRaw + buf * args.raw_buf_bytes);
          tmem_d[t] = tmem_base + acc * N;
        }
        tc_fence_after();
        const uint32_t ab0 = desc_lo(raw_addr[0], 0), ab1 = desc_lo(raw_addr[1], 0), pl16 = (uint32_t)args.plane_bytes >> 4;
        auto issue = [&](auto two) {
          constexpr bool TWO = decltype(two)::value;
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            const int ty = kb / KBX, tx = kb % KBX;
            const uint32_t shift16 = PLANES == 2 ? (uint32_t)(((ty >> 1) * SHIFT_RW + tx) * 8) + ((ty & 1) ? pl16 : 0u)
                                                 : (uint32_t)((ty * SHIFT_RW + tx) * 8);
            const uint32_t b_lo = b_lo0 + kb * (BH_BYTES >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (kb == 0 && k == 0) mma2_f16_lohi<false>(tmem_d[0], ab0 + shift16, hi0, b_lo, hi0, idesc);
              else mma2_f16_lohi<true>(tmem_d[0], ab0 + shift16 + 2 * k, hi0, b_lo + 2 * k, hi0, idesc);
            }
            if (TWO) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb == 0 && k == 0) mma2_f16_lohi<false>(tmem_d[1], ab1 + shift16, hi0, b_lo, hi0, idesc);
                else mma2_f16_lohi<true>(tmem_d[1], ab1 + shift16 + 2 * k, hi0, b_lo + 2 * k, hi0, idesc);
              }
            }
          }
          mma2_commit_both(tfull + (uint32_t)j % NACC);
          mma2_commit_both(raw_empty + (uint32_t)j % NR);     // staged pair free once every MMA reading it has retired
          if (TWO) {
            mma2_commit_both(tfull + (uint32_t)(j + 1) % NACC);
            mma2_commit_both(raw_empty + (uint32_t)(j + 1) % NR);
          }
        };
        if (elect_one()) {
          if (np == 2) issue(std::true_type{});
          else issue(std::false_type{});
        }
        __syncwarp();
      }
    }
    __syncwarp();
  } else {
    constexpr int BG = N == 64 ? 2 : 1;               // bias_mod / 32 (checked by the launcher)
    float cs[EPI == EPI_MASK_BF16 ? 32 * BG : 1];
#pragma unroll
    for (int i = 0; i < (EPI == EPI_MASK_BF16 ? 32 * BG : 1); ++i) cs[i] = 0.f;
    const int wg = warp >> 2;
    const int trow = (int)(threadIdx.x & 127);
    for (int j = 0; j < n_mine; ++j) {
      if (NEPI > 1 && (j % NEPI) != wg) continue;
      const uint32_t acc = (uint32_t)j % NACC;
      const int b = sample_of(j);
      tma_conv_epilogue_tile<N, EPI, BG>(args, b, trow, b < args.nb, tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * N, sBias, cs,
                                         tfull + acc, ((uint32_t)j / NACC) & 1, [&]() { mbar_arrive_leader_warp(tempty + acc); });
    }
    if (EPI == EPI_MASK_BF16 && args.bias_grad) {
#pragma unroll
      for (int g = 0; g < BG; ++g) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cs[g * 32 + i];
        atomicAdd(args.bias_grad + g * 32 + lane, warp_colsum32(v, lane));
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();                  // both CTAs are done with tensor memory and with each other's barriers
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// `tma_conv2s_kernel`: the CTA-pair kernel with accumulator tiles that SPAN samples.
//
// With one sample per 128-row tile a layer whose row grid has 99 rows (conv3 data gradient: 9 output rows on an
// 11-wide padded grid) uses 81 of 128 accumulator rows.  Here a CTA stages a GROUP of G samples as ONE row stream —
// sample s occupies rows [s * SROWS, (s + 1) * SROWS), SROWS = (rows per sample incl. the top padding) x RW — and the
// stream is cut into MT tiles of 128 rows (G = 5, SROWS = 99: 495 rows in 4 tiles instead of 5).  The bottom padding
// of a sample is the top padding (TMA out-of-bounds zero fill) of the next one; behind the last sample of a group
// sits a zero tail that no TMA write touches.  The boxes land at 128-byte, not 1024-byte, aligned addresses: the
// TMA unit and the MMA descriptors both apply the 128B swizzle to absolute address bits (scripts/try_unaligned_tma.py).
// Pair p processes group pairs q = p, p + pairs, ...: rank r owns samples [(2q + r) * G, +G).
// Barriers as in tma_conv2_kernel; one accumulator per tile of the group (NACC = MT).
// ---------------------------------------------------------------------------------------------------------
template <int G, int SROWS, int MT, int MAX_SHIFT>
struct TmaConv2sCfg {
  static constexpr int kBoxBytes = SROWS * 128;
  static constexpr int kZeroTailRows = MAX_SHIFT;                          // rows the last sample's windows read behind it
  static constexpr int kGroupBytes = (G * SROWS + kZeroTailRows) * 128;
  static constexpr int kSlackRows = MT * 128 + MAX_SHIFT > G * SROWS + kZeroTailRows ? MT * 128 + MAX_SHIFT - (G * SROWS + kZeroTailRows) : 0;
  static constexpr int kSlackBytes = kSlackRows * 128;                     // over-read of discarded rows behind the last group buffer
};

template <int N, int NRG, int KB, int G, int SROWS, int MT, int MAX_SHIFT>
constexpr int tma_conv2s_smem() {
  using Cf = TmaConv2sCfg<G, SROWS, MT, MAX_SHIFT>;
  return KB * (N / 2) * 128 + NRG * Cf::kGroupBytes + Cf::kSlackBytes + 1280 + 1024;
}

template <int N, int EPI, int NRG, int NEPI, int KB, int KBX, int SHIFT_RW, int G, int SROWS, int MT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(32 * (4 * NEPI + 2), 1)
    tma_conv2s_kernel(const __grid_constant__ CUtensorMap tmap, const TmaConvArgs args) {
  using namespace umma;
  constexpr int MAX_SHIFT = ((KB - 1) / KBX) * SHIFT_RW + (KBX - 1);
  using Cf = TmaConv2sCfg<G, SROWS, MT, MAX_SHIFT>;
  constexpr int BH_BYTES = (N / 2) * 128;
  constexpr int NACC = MT;
  constexpr int TMEM_COLS = NACC * N <= 64 ? 64 : NACC * N <= 128 ? 128 : NACC * N <= 256 ? 256 : 512;
  static_assert(NACC * N <= 512, "accumulators exceed tensor memory");
  static_assert(MT % 2 == 0, "tiles are issued in interleaved pairs");
  static_assert(MT * 128 >= G * SROWS, "the tiles must cover the group's row stream");
  constexpr int W_MMA = 4 * NEPI, W_TMA = 4 * NEPI + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                         // KB half weight tiles, resident
  uint8_t* sRaw = smem + KB * BH_BYTES;                       // NRG staged groups
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRaw + NRG * Cf::kGroupBytes + Cf::kSlackBytes);
  uint64_t* raw_full = bars;                    // [NRG]
  uint64_t* raw_empty = bars + NRG;             // [NRG]
  uint64_t* tfull = bars + 2 * NRG;             // [NACC]
  uint64_t* tempty = bars + 2 * NRG + NACC;     // [NACC]
  uint64_t* b_full = bars + 2 * NRG + 2 * NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NRG + 2 * NACC + 1);
  static_assert((2 * NRG + 2 * NACC + 2) * 8 <= 256, "barrier block overflows");
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int total_groups = (args.nb + G - 1) / G, total_pg = (total_groups + 1) >> 1;
  const int n_mine = pair < total_pg ? (total_pg - pair + n_pairs - 1) / n_pairs : 0;       // group pairs of this CTA pair
  auto first_of = [&](int j, int r) { return (2 * (pair + j * n_pairs) + r) * G; };         // first sample of rank r's group
  auto valid_of = [&](int j, int r) { const int v = args.nb - first_of(j, r); return v < 0 ? 0 : (v > G ? G : v); };

  // rows no TMA box writes (zero tails, sample slots of partial groups) must read as zeros / finite values
  for (int i = threadIdx.x * 16; i < NRG * Cf::kGroupBytes + Cf::kSlackBytes; i += blockDim.x * 16)
    st_shared_v4(smem_u32(sRaw + i), make_uint4(0u, 0u, 0u, 0u));
  fence_proxy_async_smem();
  if (EPI != EPI_MASK_BF16 && threadIdx.x < N) sBias[threadIdx.x] = args.bias[threadIdx.x];
  if (threadIdx.x == 0) {
    for (int a = 0; a < NRG; ++a) { mbar_init(raw_full + a, 1); mbar_init(raw_empty + a, 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 8); }
    mbar_init(b_full, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc2(tmem_slot, TMEM_COLS);
  if (warp == W_TMA && lane == 0) tma_prefetch_desc(&tmap);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto load_group = [&](int j) {      // TMA thread: this CTA's samples of group pair j into ring slot j % NRG
    const uint32_t buf = (uint32_t)j % NRG;
    if (rank == 0) mbar_arrive_expect_tx(raw_full + buf, (uint32_t)((valid_of(j, 0) + valid_of(j, 1)) * Cf::kBoxBytes));
    const int b0 = first_of(j, (int)rank), nv = valid_of(j, (int)rank);
    for (int sidx = 0; sidx < nv; ++sidx)
      tma_load_4d_2sm(sRaw + buf * Cf::kGroupBytes + sidx * Cf::kBoxBytes, &tmap, 0, -args.pad, -args.pad, b0 + sidx, raw_full + buf);
  };
  const int n_pre = n_mine < NRG ? n_mine : NRG;
  if (warp == W_TMA && lane == 0) {
    mbar_arrive_expect_tx(b_full, (uint32_t)(KB * BH_BYTES));
    for (int kb = 0; kb < KB; ++kb)
      tma_bulk_g2s(sB + kb * BH_BYTES,
                   reinterpret_cast<const uint8_t*>(args.bpack) + (int64_t)kb * (N * 128) + (int64_t)rank * BH_BYTES, BH_BYTES, b_full);
    for (int j = 0; j < n_pre; ++j) load_group(j);
  }
  __syncwarp();
  mbar_wait(b_full, 0);
  cluster_sync_all();                  // BOTH halves of the weights have landed before the leader's first MMA
  tc_fence_after();

  if (warp == W_TMA) {
    if (lane == 0) {
      for (int j = n_pre; j < n_mine; ++j) {
        mbar_wait(raw_empty + (uint32_t)j % NRG, (((uint32_t)j / NRG) & 1) ^ 1);
        load_group(j);
      }
    }
    __syncwarp();
  } else if (warp == W_MMA) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(1u, 256, N, 0, 0);
      constexpr uint32_t hi0 = desc_hi(1024, LAYOUT_SW128);
      const uint32_t b_lo0 = desc_lo(smem_u32(sB), 0);
      for (int j = 0; j < n_mine; ++j) {
        const uint32_t buf = (uint32_t)j % NRG;
        mbar_wait(raw_full + buf, ((uint32_t)j / NRG) & 1);
        const uint32_t g_addr = smem_u32(sRaw + buf * Cf::kGroupBytes);
#pragma unroll
        for (int t = 0; t < MT; t += 2) {
          mbar_wait(tempty + t, ((uint32_t)j & 1) ^ 1);
          mbar_wait(tempty + t + 1, ((uint32_t)j & 1) ^ 1);
          tc_fence_after();
          const uint32_t ab0 = desc_lo(g_addr + t * 128 * 128, 0), ab1 = desc_lo(g_addr + (t + 1) * 128 * 128, 0);
          const uint32_t d0 = tmem_base + t * N, d1 = tmem_base + (t + 1) * N;
          if (elect_one()) {
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
              const int ty = kb / KBX, tx = kb % KBX;
              const uint32_t shift16 = (uint32_t)((ty * SHIFT_RW + tx) * 8);
              const uint32_t b_lo = b_lo0 + kb * (BH_BYTES >> 4);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb == 0 && k == 0) mma2_f16_lohi<false>(d0, ab0 + shift16, hi0, b_lo, hi0, idesc);
                else mma2_f16_lohi<true>(d0, ab0 + shift16 + 2 * k, hi0, b_lo + 2 * k, hi0, idesc);
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (kb == 0 && k == 0) mma2_f16_lohi<false>(d1, ab1 + shift16, hi0, b_lo, hi0, idesc);
                else mma2_f16_lohi<true>(d1, ab1 + shift16 + 2 * k, hi0, b_lo + 2 * k, hi0, idesc);
              }
            }
            mma2_commit_both(tfull + t);
            mma2_commit_both(tfull + t + 1);
            if (t + 2 >= MT) mma2_commit_both(raw_empty + buf);     // group free once every MMA reading it has retired
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    constexpr int BG = N == 64 ? 2 : 1;
    float cs[EPI == EPI_MASK_BF16 ? 32 * BG : 1];
#pragma unroll
    for (int i = 0; i < (EPI == EPI_MASK_BF16 ? 32 * BG : 1); ++i) cs[i] = 0.f;
    const int wg = warp >> 2;
    const int trow = (int)(threadIdx.x & 127);
    for (int j = 0; j < n_mine; ++j) {
      const int b0 = first_of(j, (int)rank), nv = valid_of(j, (int)rank);
#pragma unroll
      for (int t = 0; t < MT; ++t) {
        if (NEPI > 1 && ((j * MT + t) % NEPI) != wg) continue;
        const int r = t * 128 + trow;
        const int sidx = r / SROWS, rem = r - sidx * SROWS;
        tma_conv_epilogue_tile<N, EPI, BG>(args, b0 + sidx, rem, sidx < nv, tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + t * N, sBias,
                                           cs, tfull + t, (uint32_t)j & 1, [&]() { mbar_arrive_leader_warp(tempty + t); });
      }
    }
    if (EPI == EPI_MASK_BF16 && args.bias_grad) {
#pragma unroll
      for (int g = 0; g < BG; ++g) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = cs[g * 32 + i];
        atomicAdd(args.bias_grad + g * 32 + lane, warp_colsum32(v, lane));
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

}  // namespace drl
