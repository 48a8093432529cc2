// async_copy.cuh -- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) global ->
// shared completing on an mbarrier, used by the fused kernels to stream audio
// tiles one tile ahead of the math with zero issue-slot cost on the SM.
//
//   producer (one thread):  mbar_expect_tx(bar, total_bytes);
//                           bulk_g2s(dst, src, bytes, bar);   (any number, sum == total)
//   consumers (all threads): mbar_wait(bar, parity);          parity = phase & 1
//
// Sizes and both addresses must be multiples of 16 bytes.
// Under SETK_EMU the same protocol is modelled with atomics + memcpy.
#pragma once
#include "compat.cuh"

namespace setk {

struct alignas(8) MBar {
  unsigned long long v;
};

#ifdef SETK_EMU

// layout of the emulated barrier word: [63:48] magic, [47:32] phase, [31:0] pending bytes
#define SETK_EMU_MBAR_MAGIC 0xBA55ull
__device__ inline void mbar_check(const MBar* b) {
  if ((__atomic_load_n(&b->v, __ATOMIC_SEQ_CST) >> 48) != SETK_EMU_MBAR_MAGIC)
    emu::die("mbarrier used before mbar_init (missing __syncthreads after init?)");
}
__device__ inline void mbar_init(MBar* b, int) {
  __atomic_store_n(&b->v, SETK_EMU_MBAR_MAGIC << 48, __ATOMIC_SEQ_CST);
}
__device__ inline void fence_proxy_async() {}
__device__ inline void mbar_expect_tx(MBar* b, unsigned bytes) {
  mbar_check(b);
  __atomic_fetch_add(&b->v, (unsigned long long)bytes, __ATOMIC_SEQ_CST);
}
__device__ inline void bulk_g2s(void* dst, const void* src, unsigned bytes, MBar* b) {
  mbar_check(b);
  if ((bytes & 15u) || (reinterpret_cast<uintptr_t>(dst) & 15u) || (reinterpret_cast<uintptr_t>(src) & 15u))
    emu::die("cp.async.bulk needs 16-byte aligned addresses and size");
  memcpy(dst, src, bytes);
  unsigned long long after = __atomic_sub_fetch(&b->v, (unsigned long long)bytes, __ATOMIC_SEQ_CST);
  if ((after & 0xffffffffull) == 0) __atomic_fetch_add(&b->v, 1ull << 32, __ATOMIC_SEQ_CST);  // phase++
}
__device__ inline void mbar_wait(MBar* b, unsigned parity) {
  mbar_check(b);
  while ((((__atomic_load_n(&b->v, __ATOMIC_SEQ_CST)) >> 32) & 1ull) == (unsigned long long)parity)
    std::this_thread::yield();
}
// 4-byte cp.async (LDGSTS): global -> shared without a register round trip
__device__ inline void cp_async_f32(float* dst, const float* src) { *dst = *src; }
__device__ inline void cp_async_8(void* dst, const void* src) { memcpy(dst, src, 8); }
__device__ inline void cp_async_wait_all() {}
__device__ inline void cp_async_commit() {}
__device__ inline void cp_async_wait_group1() {}

#else

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(MBar* b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// order prior generic-proxy accesses to shared memory before later async-proxy ones
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(MBar* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, MBar* b) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(b))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(MBar* b, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(b)),
      "r"(parity)
      : "memory");
}

// 4-byte cp.async (LDGSTS): global -> shared without a register round trip
__device__ __forceinline__ void cp_async_f32(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// 8-byte cp.async (one complex64 or one double)
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// all but the most recently committed group have landed
__device__ __forceinline__ void cp_async_wait_group1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

#endif

}  // namespace setk
