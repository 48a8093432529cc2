// compat.cuh -- the one place that knows whether this translation unit is
// being compiled by nvcc for sm_100a (the product) or by g++ with -DSETK_EMU
// (tests/emu: the CPU test tier's execution model for the same sources).
#pragma once

#ifdef SETK_EMU
#include "cuda_emu.h"
#define SETK_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::g_ctx->dyn)
#define SETK_HD
#else
#include <cuda_runtime.h>
#include <stdint.h>
#define SETK_DYN_SMEM(type, name) extern __shared__ __align__(1024) unsigned char name##_raw_[]; \
  type* name = reinterpret_cast<type*>(name##_raw_)
#define SETK_HD __host__ __device__
#endif

#include <atomic>

namespace setk {

extern std::atomic<long long> g_launch_count;

// Kernel launch through one funnel: counts launches (setk_launch_count) and
// lets the emulated build run the same kernel body on OS threads.
// barrier_free: the kernel uses no __syncthreads/__syncwarp/shuffles, so the
// emulator may run its threads serially.
template <class... KArgs, class... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                          void* stream, bool barrier_free, Args... args) {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
#ifdef SETK_EMU
  (void)stream;
  emu::launch(grid, block, smem, barrier_free, [&] { kernel(args...); });
  return cudaSuccess;
#else
  (void)barrier_free;
  kernel<<<grid, block, smem, static_cast<cudaStream_t>(stream)>>>(args...);
  return cudaPeekAtLastError();
#endif
}

SETK_HD inline int imin(int a, int b) { return a < b ? a : b; }
SETK_HD inline int imax(int a, int b) { return a > b ? a : b; }

}  // namespace setk
