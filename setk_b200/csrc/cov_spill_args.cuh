// cov_spill_args.cuh -- arguments of the covariance passes over the bin-major STFT workspace
// (stft_spill.cu: CUDA-core rows; cov_mma.cu: tensor-core Gram blocks) and their shared
// partial-sum layout.
#pragma once
#include "common.cuh"

namespace setk {

struct CovSpillArgs {
  const float2* xws;     // [B][T][C][pitch]
  int pitch;
  const float* mask_s; const float* mask_n; unsigned flags;
  const int* n_samples; int N; Geometry g;
  int T, F;
  int frames_per_chunk, n_chunks;
  float* partials;       // [B][n_chunks][C rows][2 masks][2*C + 1][F]
};

// floats per (utterance, chunk): rows x masks x (C complex + sum m) x F
SETK_HD inline size_t cov_spill_partial_floats(int C, int F) { return (size_t)C * 2 * (2 * C + 1) * F; }

}  // namespace setk
