// hyena-b200: Blackwell (sm_100a) tensor-core plumbing — tcgen05.mma / TMEM / mbarrier / TMA — as thin inline-PTX
// wrappers.  Used by the filter-MLP kernels (hy_filter_tc05.cu), the one place on the path that is a plain dense
// contraction (reference: src/models/sequence/hyena.py:211-219, [L,64]x[64,64] and [L,64]x[64,D]).
//
// Conventions used throughout (cta_group::1, one CTA owns its SM's tensor memory):
//   * accumulators D[128 x N] fp32 live in TMEM: lane = row (0..127), column = n; a warp w can only touch lanes
//     32*(w%4) .. 32*(w%4)+31, so "thread t of a 128-thread group" <-> row t;
//   * A comes from TMEM too (kind::tf32, one 32-bit cell per element: lane = row, column = k) — activations are
//     produced by the same thread that owns the row, so they never cross threads on their way to the tensor core;
//   * B comes from shared memory, K-major, NO swizzle: 8x(16 B) core matrices, element (n, k) at
//       (n / 8) * SBO + (k / 4) * LBO + (n % 8) * 16 + (k % 4) * 4   bytes
//     with LBO = 128 (the core matrices of one 8-row group are contiguous along K) and SBO = 128 * K/4;
//   * TF32 operands are fp32 bit patterns; 3xTF32 (hi*hi + lo*hi + hi*lo) keeps fp32-class accuracy — sin(10 x)
//     upstream amplifies rounding, one TF32 pass is not enough.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tc05 {

#define TC05_DEVICE __device__ __forceinline__

TC05_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
TC05_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
TC05_DEVICE void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TC05_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
TC05_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
TC05_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
TC05_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---- proxies / fences -------------------------------------------------------------------------------------------
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (tensor core / TMA reads)
TC05_DEVICE void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
TC05_DEVICE void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TC05_DEVICE void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one warp, all 32 lanes) ---------------------------------------------------------------------
template <int COLS>
TC05_DEVICE void tmem_alloc(uint32_t* slot_in_smem) {
  static_assert(COLS == 32 || COLS == 64 || COLS == 128 || COLS == 256 || COLS == 512, "power of two >= 32");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
TC05_DEVICE void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// TMEM address = (lane << 16) | column
TC05_DEVICE uint32_t tmem_addr(uint32_t base, int lane, int col) { return base + (static_cast<uint32_t>(lane) << 16) + col; }

// ---- TMEM <-> registers: 32 lanes x 32 bit, N consecutive columns per thread ----------------------------------------
TC05_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
TC05_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
TC05_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
TC05_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
TC05_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors ----------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor (sm_100 "version 1"), no swizzle: start address, leading (K-direction) and stride
// (M/N-direction) byte offsets of the 8 x 16 B core matrices, all in 16-byte units
TC05_DEVICE uint64_t smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
  return d;                              // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor, kind::tf32, fp32 accumulate, A and B K-major (a_major/b_major select MN-major when set)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4)                               // c_format = F32
         | (2u << 7)                             // a_format = TF32
         | (2u << 10)                            // b_format = TF32
         | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16)
         | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// ---- MMA issue (ONE thread) -------------------------------------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]^T, one K = 8 step of kind::tf32
TC05_DEVICE void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
TC05_DEVICE void mma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive (once) on the mbarrier when they have completed
TC05_DEVICE void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA (bulk tensor copy, 2-D, global -> shared, completes on an mbarrier) ------------------------------------------
TC05_DEVICE void tma_load_2d(void* smem_dst, const void* tensor_map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tensor_map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
TC05_DEVICE void tma_prefetch_desc(const void* tensor_map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tensor_map) : "memory");
}
// plain bulk copy global -> shared (no tensor map): size multiple of 16 B, both addresses 16-byte aligned
TC05_DEVICE void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- 3xTF32 split ---------------------------------------------------------------------------------------------------
// x = hi + lo exactly, hi carries the top 10 mantissa bits (round to nearest by adding half a TF32 ulp before masking:
// the tensor unit ignores the low 13 bits), lo = x - hi is exact in fp32 and is truncated by the unit (relative 2^-11 of lo)
TC05_DEVICE void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  const uint32_t h = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
  hi = h;
  lo = __float_as_uint(x - __uint_as_float(h));
}

}  // namespace tc05
