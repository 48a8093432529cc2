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

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

}  // namespace setk

// The opaque plan of the C-ABI.
struct setk_plan {
  setk_config_t cfg;
  setk::Geometry geo;
  float* d_window;       // [n_fft] analysis window, centred zero-padded, float32
  float* d_wsq;          // [n_fft] window squared (rounded from float64)
  // lazily grown workspaces (device)
  float* d_partials;     size_t partials_bytes;   // fused stft_cov partial sums
  float2* d_stft_ws;     size_t stft_ws_bytes;    // generic path STFT spill [B][C][F][T]
  float2* d_enh_ws;      size_t enh_ws_bytes;     // generic path enhanced STFT [B][F][T]
  float* d_frames_ws;    size_t frames_ws_bytes;  // generic iSTFT frames [B][T][n_fft]
  unsigned* d_peak;      size_t peak_bytes;       // [B] max|y| as uint bits
  double* d_cgmm_ws;     size_t cgmm_ws_bytes;    // CGMM posteriors, partials, R^-1 (cgmm.cu)
  int sm_count;
};
