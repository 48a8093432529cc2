// Microbenchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 up(unsigned long long r) { float2 d; asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r)); return d; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int CH>
__global__ void k_scalar(float* o, int n, float s) {
  float a[CH];
  for (int i = 0; i < CH; ++i) a[i] = threadIdx.x * 1e-3f + i;
  const float b = s, c = 1.0f - s;
  for (int j = 0; j < n; ++j)
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = fmaf(a[i], b, c);
  float r = 0; for (int i = 0; i < CH; ++i) r += a[i];
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int CH>
__global__ void k_packed(float* o, int n, float s) {
  unsigned long long a[CH];
  for (int i = 0; i < CH; ++i) a[i] = pk(threadIdx.x * 1e-3f + i, i);
  const unsigned long long b = pk(s, s * 0.5f), c = pk(1.0f - s, 0.25f);
  for (int j = 0; j < n; ++j)
#pragma unroll
    for (int i = 0; i < CH; ++i) a[i] = fma2(a[i], b, c);
  float r = 0; for (int i = 0; i < CH; ++i) { float2 v = up(a[i]); r += v.x + v.y; }
  o[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* o; cudaMalloc(&o, sizeof(float) * sms * 8 * 1024);
  const int n = 20000, blocks = sms * 8, threads = 256;      // 64 warps / SM
  constexpr int CH = 8;
  float ms_s = timeit([&] { k_scalar<CH><<<blocks, threads>>>(o, n, 0.999f); });
  float ms_p = timeit([&] { k_packed<CH><<<blocks, threads>>>(o, n, 0.999f); });
  double winstr = (double)blocks * threads / 32 * n * CH;
  printf("SMs %d clock %d kHz\n", sms, clk);
  printf("scalar FFMA : %.3f ms  %.2f warp-instr/clk/SM  %.1f TFLOP/s\n", ms_s, winstr / (ms_s * 1e-3) / (clk * 1e3) / sms, winstr * 64 / (ms_s * 1e-3) / 1e12);
  printf("packed FFMA2: %.3f ms  %.2f warp-instr/clk/SM  %.1f TFLOP/s\n", ms_p, winstr / (ms_p * 1e-3) / (clk * 1e3) / sms, winstr * 128 / (ms_p * 1e-3) / 1e12);
  return 0;
}
