// tests/emu/cuda_emu.cc -- storage for the CPU execution model's thread-locals.
// TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
#include "cuda_emu.h"
thread_local uint3 threadIdx, blockIdx;
thread_local dim3 blockDim, gridDim;
namespace emu {
thread_local Ctx* g_ctx = nullptr;
thread_local int g_tid = 0;
}
