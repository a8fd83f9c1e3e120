// umma.cuh -- thin inline-PTX layer over the sm_100a tensor-core path used by the Gram
// kernels (P3 pair counts, P5 bitmap Gram): tcgen05.mma kind::i8 with both operands in
// shared memory and int32 accumulators in tensor memory (TMEM).
//
// The Gram of a 0/1 membership matrix X (n sets x K keys, one byte per entry) is
// W = X * X^T. One tcgen05.mma consumes a K-step of 32 keys:
//   D[M x N] (+)= A[M x 32] * B[N x 32]^T     (uint8 x uint8 -> int32)
// A and B are shared-memory matrix descriptors; for the Gram both point into the same
// staging buffer.
//
// Shared-memory operand layouts without swizzle ("interleaved"), in 16-byte units:
//   K-major : ((8, m), 2) : ((1, SBO), LBO)    core matrix = 8 rows (sets) x 16 keys
//   MN-major: ((1, m), (8, k)) : ((_, SBO), (1, LBO))  core matrix = 8 keys x 16 sets
// i.e. byte address of X[set][key] is
//   K-major : (set / 8) * SBO + (set % 8) * 16 + (key / 16) * LBO + key % 16
//   MN-major: (set / 16) * SBO + set % 16 + (key / 8) * LBO + (key % 8) * 16
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmsc {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- TMEM allocation (one full warp calls these) --------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t n_cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(n_cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t n_cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(n_cols) : "memory");
}

// ---- fences -------------------------------------------------------------------------------
__device__ __forceinline__ void fence_before_thread_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_thread_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// bounded wait (a watchdog against protocol bugs): false if the phase did not complete within
// ~2^32 SM clocks (about two seconds). Bounded by the clock, not by a poll count: under a profiler,
// MPS or preemption a slow-but-correct MMA must not turn into an error.
__device__ __forceinline__ bool mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > (1ll << 32)) return false;
  return true;
}

// ---- descriptors ------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle; lbo / sbo in bytes (multiples of 16)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;
}
// instruction descriptor for kind::i8: uint8 x uint8 -> int32, dense
//   a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand
__host__ __device__ constexpr uint32_t make_idesc_u8(int M, int N, int a_mn, int b_mn) {
  return (2u << 4)                     // D format: S32
         | (0u << 7) | (0u << 10)      // A, B format: unsigned 8 bit
         | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- MMA issue (ONE thread) and completion ---------------------------------------------------
__device__ __forceinline__ void mma_u8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM -> registers: warp w reads lanes 32*(w%4) .. +31, thread = lane, 32 columns ---------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// expand 4 membership bits (low nibble of x) to 4 bytes of 0 / 1, bit i -> byte i
__device__ __forceinline__ uint32_t nibble_to_bytes(uint32_t x) { return ((x & 0xFu) * 0x00204081u) & 0x01010101u; }

}  // namespace umma
}  // namespace kmsc
