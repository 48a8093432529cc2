// tmem.cuh -- tensor memory (TMEM, 128 lanes x 512 columns x 32 bit per SM) used as a
// per-thread CONSTANT STORE by the warp-specialised FFT warps.
//
// Why: the FFT warps of stft_cov_ws.cu are register-starved (64 registers) and fetch 62
// per-thread constants -- their 32 window values and 15 complex inter-pass twiddles -- from
// shared memory on every tile: a third of their shared-memory wavefronts, on the kernel's
// busiest unit.  The constants of a thread never change during the launch, so they are
// parked once in the thread's own TMEM lane (tcgen05.st, 32x32b shape: thread t of a warp
// <-> lane 32*(warp%4) + t, consecutive columns) and read back with tcgen05.ld, which runs
// on the tensor-memory datapath and leaves the shared-memory pipe to the data.  No MMA is
// involved; TMEM is plain storage here.
//
// Protocol (PTX ISA, tcgen05): one warp allocates a power-of-two number of columns >= 32
// and publishes the base address through shared memory; tcgen05.fence::before_thread_sync /
// barrier / tcgen05.fence::after_thread_sync orders it for the other warps; a warp reaches
// only the 32 lanes of its own quadrant (warp id % 4); loads are asynchronous until
// tcgen05.wait::ld; the allocating warp frees the columns before the CTA exits.
// Under SETK_EMU the same calls act on a per-CTA array.
#pragma once
#include "compat.cuh"

namespace setk {

#ifdef SETK_EMU

__device__ inline void tmem_alloc_warp(unsigned* slot, int ncols) {
  if (ncols < 32 || ncols > 512 || (ncols & (ncols - 1))) emu::die("tcgen05.alloc: columns");
  emu::tmem_alloc(ncols);
  *slot = 0u;
}
__device__ inline void tmem_dealloc_warp(unsigned taddr, int ncols) { (void)taddr; emu::tmem_free(ncols); }
__device__ inline void tmem_fence_before_sync() {}
__device__ inline void tmem_fence_after_sync() {}
// address of (this warp's quadrant, column col) relative to the allocation
__device__ inline unsigned tmem_addr(unsigned base, int warp_in_cta, int col) {
  return base + ((unsigned)((warp_in_cta & 3) * 32) << 16) + (unsigned)col;
}
template <int NF2>
__device__ inline void tmem_st(unsigned taddr, const float2 (&r)[NF2]) {
  unsigned* row = emu::tmem_row(taddr >> 16, threadIdx.x & 31, taddr & 0xffffu, 2 * NF2);
  memcpy(row + (taddr & 0xffffu), r, sizeof(float2) * NF2);
}
__device__ inline void tmem_wait_st() {}
template <int NF2>
__device__ inline void tmem_ld(unsigned taddr, float2 (&r)[NF2]) {
  const unsigned* row = emu::tmem_row(taddr >> 16, threadIdx.x & 31, taddr & 0xffffu, 2 * NF2);
  memcpy(r, row + (taddr & 0xffffu), sizeof(float2) * NF2);
}
template <int NF2>
__device__ inline void tmem_wait_ld(float2 (&r)[NF2]) { (void)r; }

#else

__device__ __forceinline__ void tmem_alloc_warp(unsigned* slot, int ncols) {   // ONE full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   static_cast<unsigned>(__cvta_generic_to_shared(slot))),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_warp(unsigned taddr, int ncols) {  // the same warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ unsigned tmem_addr(unsigned base, int warp_in_cta, int col) {
  return base + ((unsigned)((warp_in_cta & 3) * 32) << 16) + (unsigned)col;
}

// NF2 float2 = 2*NF2 consecutive columns of the calling thread's lane (32x32b shape)
template <int NF2>
__device__ __forceinline__ void tmem_st(unsigned taddr, const float2 (&r)[NF2]);
template <>
__device__ __forceinline__ void tmem_st<4>(unsigned taddr, const float2 (&r)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "f"(r[0].x), "f"(r[0].y), "f"(r[1].x), "f"(r[1].y), "f"(r[2].x), "f"(r[2].y), "f"(r[3].x),
               "f"(r[3].y)
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int NF2>
__device__ __forceinline__ void tmem_ld(unsigned taddr, float2 (&r)[NF2]);
template <>
__device__ __forceinline__ void tmem_ld<4>(unsigned taddr, float2 (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0].x), "=f"(r[0].y), "=f"(r[1].x), "=f"(r[1].y), "=f"(r[2].x), "=f"(r[2].y),
                 "=f"(r[3].x), "=f"(r[3].y)
               : "r"(taddr)
               : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<8>(unsigned taddr, float2 (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15}, [%16];"
      : "=f"(r[0].x), "=f"(r[0].y), "=f"(r[1].x), "=f"(r[1].y), "=f"(r[2].x), "=f"(r[2].y), "=f"(r[3].x),
        "=f"(r[3].y), "=f"(r[4].x), "=f"(r[4].y), "=f"(r[5].x), "=f"(r[5].y), "=f"(r[6].x), "=f"(r[6].y),
        "=f"(r[7].x), "=f"(r[7].y)
      : "r"(taddr)
      : "memory");
}
// The loaded registers are valid only after this; they are inout operands so that the compiler
// cannot move a use above the wait.
template <int NF2>
__device__ __forceinline__ void tmem_wait_ld(float2 (&r)[NF2]);
template <>
__device__ __forceinline__ void tmem_wait_ld<4>(float2 (&r)[4]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(r[0].x), "+f"(r[0].y), "+f"(r[1].x), "+f"(r[1].y), "+f"(r[2].x), "+f"(r[2].y),
                 "+f"(r[3].x), "+f"(r[3].y)::"memory");
}
template <>
__device__ __forceinline__ void tmem_wait_ld<8>(float2 (&r)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(r[0].x), "+f"(r[0].y), "+f"(r[1].x), "+f"(r[1].y), "+f"(r[2].x), "+f"(r[2].y),
                 "+f"(r[3].x), "+f"(r[3].y), "+f"(r[4].x), "+f"(r[4].y), "+f"(r[5].x), "+f"(r[5].y),
                 "+f"(r[6].x), "+f"(r[6].y), "+f"(r[7].x), "+f"(r[7].y)::"memory");
}

#endif

}  // namespace setk
