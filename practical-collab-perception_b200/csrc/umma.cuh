// umma.cuh - hand-written sm_100a tensor-core plumbing: tcgen05.mma (kind::tf32) issued by one thread,
// operands in shared memory described by 64-bit matrix descriptors, accumulators in TMEM, completion
// through an mbarrier, results read back with tcgen05.ld.  No CUTLASS/CuTe: inline PTX only.
//
// Operand layout used throughout (K-major, SWIZZLE_NONE "interleaved" canonical layout):
//   a [rows x K] fp32 operand is stored as K/4 panels; panel kc holds columns 4kc..4kc+3 of every row:
//       byte address = base + kc * (rows * 16) + row * 16 + (k % 4) * 4
//   i.e. 8 rows x 16 bytes form one contiguous 128-byte core matrix,
//       SBO (next 8-row group)      = 128 bytes
//       LBO (next 16-byte K chunk)  = rows * 16 bytes
//   One tcgen05.mma of kind::tf32 consumes K = 8 (two panels); the next K step starts 2 panels further.
//   A thread that owns a row writes one float4 per panel; consecutive lanes hit consecutive 16-byte
//   slots, so the stores are bank-conflict free.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcp {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// exactly one lane of a converged warp returns true.  Issuing tcgen05.mma / tcgen05.commit under this predicate (and
// not under `threadIdx.x == 0`) lets the compiler emit the uniform-datapath instructions back to back; a thread-id
// test makes it wrap every UTCHMMA in an ELECT / BRA.U.ANY uniformisation loop (~60 cycles per MMA, measured).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// mbarrier.try_wait suspends the thread in hardware until the phase completes or an implementation-defined time limit
// expires.  PCP_MBAR_HINT > 0 passes an explicit suspend-time hint (ns): ptxas then adds a NANOSLEEP.SYNCS per failed
// attempt, which wakes on every arrival at the barrier - with 256-arrival barriers that is hundreds of re-polls per slot.
#ifndef PCP_MBAR_HINT
#define PCP_MBAR_HINT 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
#if PCP_MBAR_HINT > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)PCP_MBAR_HINT)
      : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
#endif
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- proxies / fences -----------------------------------------------------------------------------
// generic-proxy st.shared -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp executes these) -----------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------
// 64-bit shared-memory matrix descriptor, K-major, no swizzle (layout documented at the top of this file)
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t rows) {
  const uint32_t lbo = rows * 16u, sbo = 128u;
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version: Blackwell
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// 32-bit instruction descriptor: D = fp32, A = B = tf32, both K-major, M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_tf32_m128(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[128 x N] (+)= A[128 x 8] . B[N x 8]^T   (one K step), issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// same with the A operand in TENSOR MEMORY: A[128 x 8] occupies lanes 0..127 x 8 consecutive 32-bit columns at a_tmem
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 3xTF32: D (+)= (A_hi + A_lo) . (B_hi + B_lo)^T without the lo.lo term, over `ksteps` K steps of 8.
// a_hi/a_lo: operand bases with `a_rows` rows (128); b_hi/b_lo with `b_rows` rows (= N).
__device__ __forceinline__ void mma_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t a_rows, uint32_t b_hi,
                                           uint32_t b_lo, uint32_t b_rows, int ksteps, uint32_t idesc, bool accumulate) {
  for (int s = 0; s < ksteps; ++s) {
    const uint32_t ao = (uint32_t)s * 2u * a_rows * 16u, bo = (uint32_t)s * 2u * b_rows * 16u;
    const uint64_t ah = smem_desc_kmajor(a_hi + ao, a_rows), al = smem_desc_kmajor(a_lo + ao, a_rows);
    const uint64_t bh = smem_desc_kmajor(b_hi + bo, b_rows), bl = smem_desc_kmajor(b_lo + bo, b_rows);
    mma_tf32(d_tmem, al, bh, idesc, accumulate || s > 0);   // small terms first
    mma_tf32(d_tmem, ah, bl, idesc, true);
    mma_tf32(d_tmem, ah, bh, idesc, true);
  }
}

// 3xTF32 with A (hi at a_hi_tmem, lo at a_lo_tmem, K consecutive columns each) in tensor memory
__device__ __forceinline__ void mma_3xtf32_ts(uint32_t d_tmem, uint32_t a_hi_tmem, uint32_t a_lo_tmem, uint32_t b_hi,
                                              uint32_t b_lo, uint32_t b_rows, int ksteps, uint32_t idesc, bool accumulate) {
  for (int s = 0; s < ksteps; ++s) {
    const uint32_t bo = (uint32_t)s * 2u * b_rows * 16u;
    const uint64_t bh = smem_desc_kmajor(b_hi + bo, b_rows), bl = smem_desc_kmajor(b_lo + bo, b_rows);
    mma_tf32_ts(d_tmem, a_lo_tmem + 8u * s, bh, idesc, accumulate || s > 0);
    mma_tf32_ts(d_tmem, a_hi_tmem + 8u * s, bl, idesc, true);
    mma_tf32_ts(d_tmem, a_hi_tmem + 8u * s, bh, idesc, true);
  }
}

// ---- registers -> TMEM: this warp's 32 lanes x 16 consecutive 32-bit columns ---------------------
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// tcgen05.ld without the wait (several loads in flight, one wait)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t& r4,
                                                uint32_t& r5, uint32_t& r6, uint32_t& r7) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- TMEM -> registers: this warp's 32 lanes x 16 consecutive fp32 columns -----------------------
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- cp.async (LDGSTS): 4-byte global -> shared copies that need no registers -------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// ---- bulk (TMA) copies shared -> global: one thread sends `bytes` (multiple of 16) as whole lines ----
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TF32 split ------------------------------------------------------------------------------------
// x = hi + lo exactly, hi has a 10-bit mantissa (truncation); the tensor core drops the low 13 bits of lo.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = __fsub_rn(x, hi);
}
// round-to-nearest variant for the (pre-packed) weights
__device__ __forceinline__ void split_tf32_rn(float x, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  lo = __fsub_rn(x, hi);
}

}  // namespace umma
}  // namespace pcp
