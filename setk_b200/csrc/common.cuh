// common.cuh -- plan object, frame bookkeeping and small complex helpers shared
// by every kernel of libsetk_b200.
#pragma once
#include "compat.cuh"
#include "../../include/setk_b200.h"

namespace setk {

// float32 machine epsilon: EPSILON of the reference (libs/utils.py:16)
#define SETK_EPS32 1.1920928955078125e-07f
#define SETK_EPS32_D 1.1920928955078125e-07
// numpy.finfo(float32).tiny: librosa.util.tiny for a float32 iSTFT buffer
#define SETK_TINY32 1.17549435e-38f

// internal flag (not part of the C-ABI): use 1 - mask (after clipping)
#define SETK_F_ONE_MINUS_INTERNAL 0x80000000u

struct Geometry {   // passed by value to kernels
  int C;            // channels
  int n_fft;        // FFT size (power of two)
  int log2n;
  int hop;
  int pad;          // n_fft/2 if center else 0
  int F;            // n_fft/2 + 1
};

// T of librosa.stft for an utterance of n samples;
// <= 0 when the utterance is too short for one frame / for reflect padding.
SETK_HD inline int frames_of(int n, int n_fft, int hop, int pad) {
  if (pad > 0 && n < pad + 1) return 0;   // np.pad(reflect) needs n > pad
  int padded = n + 2 * pad;
  if (padded < n_fft) return 0;
  return 1 + (padded - n_fft) / hop;
}

// index into the un-padded signal of padded position p (np.pad mode="reflect")
SETK_HD inline int reflect_index(int p, int pad, int n) {
  int i = p - pad;
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

struct cf { float x, y; };  // not used for storage; float2 is

// ---------------------------------------------------------------------------
// Packed fp32 pairs.  sm_100a executes add/sub/mul/fma.f32x2 (SASS FADD2 /
// FMUL2 / FFMA2) on 64-bit register pairs with per-operand half swizzles,
// per-half negation and scalar broadcast, so ptxas folds the (re, im) shuffles
// written below into the instruction: a complex add is ONE instruction, a
// complex multiply TWO.  The fp32 pipe does the same flops per clock either way
// (measured: tools/micro/ffma2_rate.cu, 72 vs 64 TFLOP/s); what halves is the
// number of issue slots, which is what bounds the STFT kernels.
// ---------------------------------------------------------------------------
#ifdef SETK_EMU
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
}
#else
__device__ __forceinline__ unsigned long long f2pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2unpack(unsigned long long r) {
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r));
  return d;
}
__device__ __forceinline__ float2 f2add(float2 a, float2 b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2pack(a)), "l"(f2pack(b)));
  return f2unpack(d);
}
__device__ __forceinline__ float2 f2sub(float2 a, float2 b) {
  unsigned long long d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2pack(a)), "l"(f2pack(b)));
  return f2unpack(d);
}
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(f2pack(a)), "l"(f2pack(b)));
  return f2unpack(d);
}
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(f2pack(a)), "l"(f2pack(b)), "l"(f2pack(c)));
  return f2unpack(d);
}
#endif

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return f2add(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return f2sub(a, b); }
// a * b = (ax bx, ay bx) + (-ay by, ax by)
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return f2fma(make_float2(a.y, a.x), make_float2(-b.y, b.y), f2mul(a, make_float2(b.x, b.x)));
}
// a * b + c
__device__ __forceinline__ float2 cmad(float2 a, float2 b, float2 c) {
  return f2fma(make_float2(a.y, a.x), make_float2(-b.y, b.y), f2fma(a, make_float2(b.x, b.x), c));
}
// a * conj(b) = (ax bx, ay bx) + (ay by, -ax by)
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {
  return f2fma(make_float2(a.y, a.x), make_float2(b.y, -b.y), f2mul(a, make_float2(b.x, b.x)));
}
// conj(w) * x + c = (wx xx, wx xy) + (wy xy, -wy xx) + c
__device__ __forceinline__ float2 cmad_conjw(float2 w, float2 x, float2 c) {
  return f2fma(make_float2(x.y, x.x), make_float2(w.y, -w.y), f2fma(x, make_float2(w.x, w.x), c));
}

}  // namespace setk

// The opaque plan of the C-ABI.
struct setk_plan {
  setk_config_t cfg;
  setk::Geometry geo;
  float* d_window;       // [n_fft] analysis window, centred zero-padded, float32
  float* d_wsq;          // [n_fft] window squared (rounded from float64)
  float win_pair_sum;    // K when window[n] + window[n + n_fft/2] == K for every n (Hann: 1), else 0
  // lazily grown workspaces (device)
  float* d_partials;     size_t partials_bytes;   // fused stft_cov partial sums
  float2* d_stft_ws;     size_t stft_ws_bytes;    // generic path STFT spill [B][C][F][T]
  float2* d_enh_ws;      size_t enh_ws_bytes;     // generic path enhanced STFT [B][F][T]
  float* d_frames_ws;    size_t frames_ws_bytes;  // generic iSTFT frames [B][T][n_fft]
  unsigned* d_peak;      size_t peak_bytes;       // [B] max|y| as uint bits
  double* d_cgmm_ws;     size_t cgmm_ws_bytes;    // CGMM posteriors, partials, R^-1 (cgmm.cu)
  int* d_tile_prefix;   size_t tile_prefix_bytes;   // ragged batches: tiles before each utterance
  float* d_wave_ws;     size_t wave_ws_bytes;       // un-normalised wave of the PCM-16 output path
  int sm_count;
};
