// stft_cov_args.cuh -- launch arguments and accumulator bookkeeping shared by the two
// fused STFT + covariance kernels (stft_cov_fused.cu, stft_cov_ws.cu).
#pragma once
#include "common.cuh"
#include "stft_tile.cuh"

namespace setk {

template <int C>
struct CovAcc {
  static constexpr int NOFF = C * (C - 1) / 2;
  static constexpr int NACC = C * C;  // C real diagonals + NOFF complex
};

struct StftCovArgs {
  Geometry g;
  const float* audio; const int* n_samples; int N;
  const float* mask_s; const float* mask_n; unsigned flags;
  int T;                 // frames of an N-sample utterance (mask leading dim)
  TileSched sched;       // which (utterance, tile) pairs this CTA owns
  int slots;             // partial-sum slots per utterance
  const float* window;   // [n_fft]
  float win_pair_sum;    // K if window[n] + window[n + n_fft/2] == K for all n (Hann: 1), else 0
  float* partials;       // [B][slots][2*C*C + 2][F]
  unsigned* maxabs_bits; // [B] or null
};

}  // namespace setk
