// Family 1d — `tma_gemm2_kernel`: the fc-layer GEMMs on CTA PAIRS (tcgen05 `cta_group::2`).
//
// Why: a 128 x 256 tile moves 48 KB per k-block through the SM's shared-memory port twice (TMA write + operand
// fetch) for 512 tensor cycles = 192 B/clk against a 128 B/clk port, so the single-CTA kernel holds the tensor
// pipe at ~60 %.  A CTA pair computes a 256 x 256 tile: each CTA stages ITS 128 rows of A and HALF of B
// (128 of the 256 columns), the pair's MMA (issued by the leader CTA only, M = 256) reads both halves of B, and
// each SM's port carries 32 KB per k-block twice = 128 B/clk.
//
// Cluster (2,1,1), 320 threads per CTA: warps 0-7 two epilogue warpgroups (each CTA drains its own 128
// accumulator rows from its own TMEM), warp 8 MMA (leader CTA issues; both CTAs allocate TMEM with cta_group::2),
// warp 9 TMA (both CTAs load their halves; the transaction bytes of BOTH land on the leader's `full` barrier).
//   full[s]   leader only: 1 arrival (expect_tx of 64 KB) + the four loads of the pair
//   empty[s]  both CTAs: tcgen05.commit multicast after the MMAs that read stage s
//   tfull[a]  both CTAs: tcgen05.commit multicast after the last MMA of a tile
//   tempty[a] leader only: 512 arrivals = both CTAs' epilogue threads (remote arrive from the peer)
#pragma once
#include <cuda.h>
#include "umma_tma_gemm.cuh"

namespace drl {

constexpr int kG2Stages = 6;
constexpr int kG2Stage = 2 * 128 * 128;                 // A 128 x 64 + B half 128 x 64, bf16
constexpr int kG2Smem = kG2Stages * kG2Stage + 256 + 1024 + 1024;
constexpr int kG2Threads = 320;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA tile load whose completion bytes are credited to the LEADER CTA's mbarrier (same smem offset, CTA-rank bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          umma::smem_u32(smem_dst)),
      "l"(tmap), "r"(umma::smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
template <bool ACC>
__device__ __forceinline__ void mma2_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc) {
  if (ACC) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  }
}
// all MMAs issued so far by this thread -> one arrival on `bar` in BOTH CTAs of the pair
__device__ __forceinline__ void mma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   umma::smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same smem offset in the leader CTA (rank 0)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(umma::smem_u32(bar))
      : "memory");
}

// The same for a whole converged warp with ONE remote arrival: 128 remote arrivals per accumulator tile serialise on the
// leader's barrier (the CTA-pair conv kernels sat on a 0.20 ms floor of these handshakes); barrier counts are per warp.
__device__ __forceinline__ void mbar_arrive_leader_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive_leader(bar);
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
    tma_gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const TmaGemmArgs args) {
  using namespace umma;
  constexpr int BN = 256, S = kG2Stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * kG2Stage);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint64_t* tempty = bars + 2 * S + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  float* sBias = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int m_pairs = (args.M + 255) / 256, n_tiles = (args.N_total + BN - 1) / BN;
  const int num_tiles = m_pairs * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    // tempty: 16 arrivals = 2 CTAs x 2 epilogue warpgroups x 4 warps (mbar_arrive_leader_warp)
    for (int a = 0; a < 2; ++a) { mbar_init(tfull + a, 1); mbar_init(tempty + a, 16); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc2(tmem_slot, 512);
  if (warp == 9 && lane == 0) { tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); }
  tc_fence_before();
  cluster_sync_all();                  // barriers of both CTAs initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs) {
        const int nt = tile % n_tiles, mp = tile / n_tiles;     // N tiles of one row block adjacent: pairs p, p+1 fetch the same A rows together (one DRAM read, one L2 hit)
        for (int kb = 0; kb < args.KB; ++kb, ++it) {
          const uint32_t stage = it % S, phase = (it / S) & 1;
          mbar_wait(empty + stage, phase ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(full + stage, 2 * kG2Stage);
          uint8_t* st = smem + stage * kG2Stage;
          tma_load_2d_2sm(st, &tmap_a, kb * 64, mp * 256 + (int)rank * 128, full + stage);
          tma_load_2d_2sm(st + 128 * 128, &tmap_b, kb * 64, nt * BN + (int)rank * 128, full + stage);
        }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_f16(1u, 256, BN, 0, 0);
      constexpr uint32_t hi = desc_hi(1024, LAYOUT_SW128);
      const uint32_t a_lo0 = desc_lo(smem_u32(smem), 0), b_lo0 = desc_lo(smem_u32(smem) + 128 * 128, 0);
      uint32_t it = 0, j = 0;
      for (int tile = pair; tile < num_tiles; tile += n_pairs, ++j) {
        const uint32_t acc = j & 1;
        mbar_wait(tempty + acc, ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < args.KB; ++kb, ++it) {
          const uint32_t stage = it % S, phase = (it / S) & 1;
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (kG2Stage >> 4), b_lo = b_lo0 + stage * (kG2Stage >> 4);
          if (elect_one()) {
            if (kb == 0) mma2_f16_lohi<false>(tmem_d, a_lo, hi, b_lo, hi, idesc);
            else mma2_f16_lohi<true>(tmem_d, a_lo, hi, b_lo, hi, idesc);
            mma2_f16_lohi<true>(tmem_d, a_lo + 2, hi, b_lo + 2, hi, idesc);
            mma2_f16_lohi<true>(tmem_d, a_lo + 4, hi, b_lo + 4, hi, idesc);
            mma2_f16_lohi<true>(tmem_d, a_lo + 6, hi, b_lo + 6, hi, idesc);
            mma2_commit_both(empty + stage);
            if (kb == args.KB - 1) mma2_commit_both(tfull + acc);
          }
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    uint32_t j = 0;
    int last_nt = -1;
    constexpr int BG = 2;
    float cs[EPI == EPI_MASK_BF16 ? 32 * BG : 1];
#pragma unroll
    for (int i = 0; i < (EPI == EPI_MASK_BF16 ? 32 * BG : 1); ++i) cs[i] = 0.f;
    for (int tile = pair; tile < num_tiles; tile += n_pairs, ++j) {
      const int nt = tile % n_tiles, mp = tile / n_tiles;
      const uint32_t acc = j & 1;
      const int m = mp * 256 + (int)rank * 128 + (int)(threadIdx.x & 127);
      const int n0 = nt * BN;
      if (!epi_is_mask(EPI) && nt != last_nt) {            // (re)load the bias slice of this N tile
        asm volatile("bar.sync 1, 256;");
        for (int i = threadIdx.x; i < BN; i += 256) sBias[i] = n0 + i < args.N_total ? args.bias[n0 + i] : 0.f;
        asm volatile("bar.sync 1, 256;");
        last_nt = nt;
      }
      mbar_wait(tfull + acc, (j >> 1) & 1);
      tc_fence_after();
      const int wg = warp >> 2;
#pragma unroll
      for (int c2 = 0; c2 < BN / 64; ++c2) {
        const int ch = wg * (BN / 64) + c2;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + acc * BN + ch * 32, r);
        tmem_ld_wait();
        if (c2 == BN / 64 - 1) {
          tc_fence_before();
          mbar_arrive_leader_warp(tempty + acc);
        }
        const int n = n0 + ch * 32;
        if (m < args.M && n < args.N_total) {
          if (EPI == EPI_BIAS_RELU_F32) {
            float* o = reinterpret_cast<float*>(args.out) + (int64_t)m * args.out_ld + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(r[8 * q + i]) + sBias[ch * 32 + 8 * q + i], 0.f));
              stg_v8(o + 8 * q, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
            }
          } else if (EPI == EPI_BIAS_RELU_BF16) {        // bf16 features for the tensor-core heads kernel (umma_heads.cuh)
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + (int64_t)m * args.out_ld + n;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i)
                pk[i] = pack_bf16x2_relu(__uint_as_float(r[16 * q + 2 * i]) + sBias[ch * 32 + 16 * q + 2 * i],
                                         __uint_as_float(r[16 * q + 2 * i + 1]) + sBias[ch * 32 + 16 * q + 2 * i + 1]);
              stg_v8(o + 16 * q, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            }
          } else {
            const int64_t off = (int64_t)m * args.out_ld + n;
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(args.out) + off;
            uint32_t mb;
            if (args.bits_in) {
              mb = __ldg(args.bits_in + (off >> 5));
            } else {
              uint32_t mw[16];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 mv = ldg_v4(args.mask_src + off + 8 * q);
                mw[4 * q] = mv.x; mw[4 * q + 1] = mv.y; mw[4 * q + 2] = mv.z; mw[4 * q + 3] = mv.w;
              }
              mb = relu_bits_from_pairs(mw);
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float lo = (mb >> relu_bit_pos(16 * q + 2 * i)) & 1u ? __uint_as_float(r[16 * q + 2 * i]) : 0.f;
                const float hi2 = (mb >> relu_bit_pos(16 * q + 2 * i + 1)) & 1u ? __uint_as_float(r[16 * q + 2 * i + 1]) : 0.f;
                pk[i] = pack_bf16x2(lo, hi2);
                r[16 * q + 2 * i] = __float_as_uint(lo);
                r[16 * q + 2 * i + 1] = __float_as_uint(hi2);
              }
              stg_v8(o + 16 * q, pk[0], pk[1], pk[2], pk[3], pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        if (EPI == EPI_MASK_BF16 && m < args.M && n < args.N_total) {
#pragma unroll
          for (int i = 0; i < 32; ++i) cs[(c2 % BG) * 32 + i] += __uint_as_float(r[i]);
        }
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
  cluster_sync_all();                  // both CTAs are done with tensor memory and with each other's barriers
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

}  // namespace drl
