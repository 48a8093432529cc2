// apply_istft_args.cuh -- launch arguments shared by the two fused apply + iSTFT kernels
// (apply_istft_fused.cu).
#pragma once
#include "common.cuh"
#include "stft_tile.cuh"

namespace setk {

struct ApplyIstftArgs {
  Geometry g;
  const float* audio; const int* n_samples; int N;
  const void* w; int w_dtype;
  const float* post_mask; int T;   // mask leading dimension (frames of N samples)
  TileSched sched;      // which (utterance, tile) pairs this CTA owns
  const float* window;  // [n_fft] analysis == synthesis window
  const float* wsq;     // [n_fft]
  int n_out;
  float* wave;          // [B][n_out]
  unsigned* peak;       // [B] or null
  int c0, c_total;      // this launch handles channels [c0, c0 + C) of c_total
  int accumulate;       // != 0: wave += this block's contribution (iSTFT is linear)
};

}  // namespace setk
