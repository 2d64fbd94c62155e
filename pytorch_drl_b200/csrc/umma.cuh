// Hand-written sm_100a primitives: mbarrier, 1-D TMA bulk copies, tcgen05 (alloc / mma / commit /
// ld), shared-memory matrix descriptors and the 128B/64B swizzled tile layouts they describe.
// No CUTLASS/CuTe: everything is inline PTX.  Bit layouts follow the PTX ISA "tcgen05 matrix
// descriptor" / "instruction descriptor" tables.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace drl {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// One arrival for a whole CONVERGED warp (barrier counts are then per warp): 32 arrivals on one shared-memory word serialise,
// and every one of them is a wavefront on the shared-memory data pipe the tensor cores fetch their operands through.
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
// try_wait suspends the thread in hardware until the phase completes or the time hint (ns) expires,
// so waiting warps do not burn the issue slots of the producer warp that shares their scheduler.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// generic-proxy smem writes -> visible to the async proxy (tensor core / TMA) of this CTA
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- 1-D TMA bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP) -----------------
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- tcgen05: TMEM allocation ------------------------------------------------------------------------
// Must be executed by one full warp.  ncols: power of two in [32, 512].
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------------
enum : uint64_t { LAYOUT_SW128 = 2, LAYOUT_SW64 = 4, LAYOUT_SW32 = 6, LAYOUT_NONE = 0 };

// Shared-memory matrix descriptor: start address, leading / stride byte offsets (all >> 4),
// descriptor version 1 (bit 46), swizzle mode in bits [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint64_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}

// Instruction descriptor for kind::f16 with fp32 accumulation.
//   fmt: 0 = fp16, 1 = bf16 (both operands);  a_mn / b_mn: 1 = MN-major operand, 0 = K-major.
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t fmt, uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Issued by ONE thread: call inside `if (elect_one()) { ... }` — see elect_one() below.
// Split descriptor form for the issue loop: the high word (SBO, version, swizzle) is a constant of the
// kernel, the low word is (address >> 4) | LBO << 16 and advances by plain 32-bit adds, so one MMA
// costs ~4 instructions of the single issuing thread instead of ~20.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes, uint32_t layout) {
  return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
}
template <bool ACC>
__device__ __forceinline__ void mma_f16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc) {
  if (ACC) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  }
}

// Integer form: u8 / s8 operands (formats in the instruction descriptor), s32 accumulators, K = 32 bytes per instruction.
template <bool ACC>
__device__ __forceinline__ void mma_i8_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc) {
  if (ACC) {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .b64 da, db;\n\t.reg .pred p;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  }
}

// D[tmem] (+)= A[tmem] * B[smem]: the A operand (K-major only) is read from tensor memory — lane = tile row,
// one 32-bit column = two consecutive 16-bit K elements, 8 columns per K=16 instruction.
template <bool ACC>
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  if (ACC) {
    asm volatile(
        "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .b64 db;\n\t.reg .pred p;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc)
        : "memory");
  }
}

// registers -> TMEM: this warp's 32 lanes x 32 consecutive columns (thread t writes lane 32*(warp%4)+t).
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// One lane of the (converged) warp.  MMA-issuing warps keep ALL lanes in the loop (waits, address math) and wrap
// each unrolled BATCH of tcgen05 instructions in `if (elect_one()) { ... }`: addresses and descriptors then stay warp-uniform for the compiler
// (uniform registers feed UTCHMMA directly); a loop run under `if (lane == 0)` instead makes every operand a
// per-thread value and costs an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall per MMA.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// All previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive columns (one fp32 row chunk per thread).
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- swizzled tile addressing -----------------------------------------------------------------------------
// A tile is a stack of rows of ROWB bytes (128 or 64); 16-byte chunk c of row r lives at
//   r*ROWB + ((c ^ f(r)) << 4),  f(r) = r & 7 (128B swizzle) or (r >> 1) & 3 (64B swizzle),
// i.e. Swizzle<3,4,3> / Swizzle<2,4,3> of the byte offset.  Tile bases are 1024-byte aligned.
template <int ROWB>
__device__ __forceinline__ uint32_t swz_off(uint32_t r, uint32_t c) {
  static_assert(ROWB == 128 || ROWB == 64, "row bytes");
  if (ROWB == 128) return r * 128u + ((c ^ (r & 7u)) << 4);
  return r * 64u + ((c ^ ((r >> 1) & 3u)) << 4);
}

__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 32-byte global store (sm_100: STG.E.256): one full 32-byte sector per thread and instruction.  Row epilogues
// that write 16 bytes per thread and instruction send two half-sector requests per sector to L2 (measured: the
// L1->L2 write traffic of every row epilogue was exactly 2x the tensor size).
__device__ __forceinline__ void stg_v8(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e, uint32_t f,
                                       uint32_t g, uint32_t h) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "r"(e),
               "r"(f), "r"(g), "r"(h)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_v2(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

// ---- Ampere-style async copies (LDGSTS): 16 B global -> shared without register staging ---------------
// src_bytes = 0 zero-fills the destination (padding taps / rows past the end).
__device__ __forceinline__ void cp_async_16(uint32_t saddr, const void* gptr, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(gptr), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 8 uint8 (two 32-bit words) -> 8 fp16, exact: byte b becomes the half 0x6400|b = 1024+b, then -1024.
__device__ __forceinline__ uint4 u8x8_to_f16x8(uint2 w) {
  const uint32_t magic = 0x64006400u;
  uint32_t h0 = __byte_perm(w.x, magic, 0x7170);   // bytes {b0, 0x64, b1, 0x64}
  uint32_t h1 = __byte_perm(w.x, magic, 0x7372);   // {b2, 0x64, b3, 0x64}
  uint32_t h2 = __byte_perm(w.y, magic, 0x7170);
  uint32_t h3 = __byte_perm(w.y, magic, 0x7372);
  const __half2 k = __half2half2(__ushort_as_half((unsigned short)0x6400));
  uint4 o;
  __half2 t;
  t = __hsub2(*reinterpret_cast<__half2*>(&h0), k); o.x = *reinterpret_cast<uint32_t*>(&t);
  t = __hsub2(*reinterpret_cast<__half2*>(&h1), k); o.y = *reinterpret_cast<uint32_t*>(&t);
  t = __hsub2(*reinterpret_cast<__half2*>(&h2), k); o.z = *reinterpret_cast<uint32_t*>(&t);
  t = __hsub2(*reinterpret_cast<__half2*>(&h3), k); o.w = *reinterpret_cast<uint32_t*>(&t);
  return o;
}

// Column sums across a warp: every lane holds one row of 32 values; returns in lane j the sum over the
// 32 lanes of v[j] (a transpose-reduce: 31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// ---- ReLU mask words ------------------------------------------------------------------------------------------
// One 32-bit word per 32 consecutive output elements (word index = element index >> 5).  The bit of
// element k (0..31) inside its word sits at relu_bit_pos(k): elements are produced as bf16x2 pairs, pair i =
// elements (2i, 2i+1), and  (pair + 0x7FFF7FFF)  raises bit 15 / bit 31 exactly when the low / high bf16 (both
// non-negative after the ReLU) is non-zero, so the producer needs 3 integer ops per PAIR:
//     acc = (acc >> 1) | ((pair_i + 0x7FFF7FFF) & 0x80008000)      for i = 15 .. 0
// which leaves element 2i at bit 15 - i and element 2i + 1 at bit 31 - i.
__host__ __device__ constexpr int relu_bit_pos(int k) { return ((k & 1) ? 31 : 15) - (k >> 1); }
__device__ __forceinline__ uint32_t relu_bits_from_pairs(const uint32_t (&w)[16]) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 15; i >= 0; --i) acc = (acc >> 1) | ((w[i] + 0x7FFF7FFFu) & 0x80008000u);
  return acc;
}
// LeakyReLU outputs keep their sign: bit = (value > 0), same bit positions.  (mag + 0x7FFF) raises bit 15 / 31 iff the
// magnitude is non-zero (no carry between the halves: 0x7FFF + 0x7FFF < 0x10000); & ~w clears it for negative values.
__device__ __forceinline__ uint32_t pos_bits_from_pairs(const uint32_t (&w)[16]) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 15; i >= 0; --i) acc = (acc >> 1) | ((((w[i] & 0x7FFF7FFFu) + 0x7FFF7FFFu) & ~w[i]) & 0x80008000u);
  return acc;
}
// {lo, hi} -> bf16x2 with the ReLU folded into the conversion (negative -> +0)
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

}  // namespace umma
// 16-byte vector reduction (SASS RED.E.ADD.F32x4...): one L2 atomic operation for four consecutive floats
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

}  // namespace drl
