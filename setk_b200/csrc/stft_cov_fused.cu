// stft_cov_fused.cu -- THE METRIC KERNEL: multichannel STFT fused with the
// mask-weighted spatial covariance; the STFT never touches HBM.
//
// Replaces (scripts/sptk): SpectrogramReader._load (libs/data_handler.py:492-503,
// forward_stft per channel, libs/utils.py:96-138 -> librosa.stft) followed by
// SupervisedBeamformer.run's two compute_covar calls (libs/beamformer.py:279-281,
// 87-103); C++ twin: ShortTimeFTComputer::Compute (include/stft.cc:28-66) +
// EstimatePsd (include/beamformer.cc:91-120).
//
// Mapping (template <C, 512-point frames, TT frames per tile>), 288 threads:
//   grid = (chunks, B): a CTA owns a run of frames of one utterance.
//   per tile of TT frames
//     stage   (TT-1)*hop + n_fft samples x C channels -> smem, 16-byte loads;
//             reflect padding / ragged ends resolved here; running max|x|
//     FFT     one half-warp per (frame, channel): 512-point real FFT as a
//             256-point complex FFT, 16 values per lane in registers, one
//             conflict-free smem exchange (fft16.cuh); Z stays in smem
//     cov     one thread per bin k (thread 256 = Nyquist): split Z -> X_c[k]
//             with the thread-constant twiddle, accumulate the Hermitian upper
//             triangle of  sum m x x^H  for (m_s, m_n) in fp32 registers
//   end: partial sums -> workspace [B][chunk][acc][F]; cov_finalize_kernel
//   reduces chunks in fixed order, normalises by max(sum m, 1e-6), writes
//   Rs, Rn c64 [B][F][C][C].
// Algorithmic bytes per utterance: 4*C*N + 4*T*F (+4*T*F with mask_n) + 2*8*F*C^2.
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
  int frames_per_chunk;  // multiple of TT
  int n_chunks;
  const float* window;   // [n_fft]
  float* partials;       // [B][n_chunks][2*C*C + 2][F]
  unsigned* maxabs_bits; // [B] or null
};

template <int C, int TT>
__global__ void __maxnreg__(112) stft_cov_kernel(StftCovArgs a) {
  constexpr int NACC = CovAcc<C>::NACC;
  constexpr int F = kBins;
  SETK_DYN_SMEM(float, smem);
  const int hop = a.g.hop, pad = a.g.pad;
  TileSmem<C, TT> sm;
  sm.carve(smem, hop);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int nb = a.n_samples ? a.n_samples[b] : a.N;
  const int Tb = frames_of(nb, kNfft, hop, pad);
  const int t_begin = chunk * a.frames_per_chunk;
  const int t_end = imin(t_begin + a.frames_per_chunk, Tb);

  for (int n = tid; n < kNfft; n += blockDim.x) sm.win[n] = 0.5f * a.window[n];
  if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); }

  // thread constants
  float w1s, w1c;
  sincospif((float)(lane & 15) / 128.0f, &w1s, &w1c);      // W256^{lane16}
  const float2 w1 = make_float2(w1c, -w1s);
  const int bin = tid;                                      // cov role: bins 0..256
  const bool cov_thread = bin < F;
  const float2 tw = split_twiddle(bin);
  const int zk = bin & (kM - 1), zn = (kM - bin) & (kM - 1);

  // accumulators as packed pairs (FFMA2): dg[i] = (sum m_s |x_i|^2, sum m_n |x_i|^2),
  // os[p] / on[p] = sum m x_i conj(x_k) for the p-th (i < k), sm = (sum m_s, sum m_n)
  constexpr int NOFF = C * (C - 1) / 2;
  float2 dg[C], os[NOFF > 0 ? NOFF : 1], on[NOFF > 0 ? NOFF : 1];
#pragma unroll
  for (int i = 0; i < C; ++i) dg[i] = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NOFF; ++i) { os[i] = make_float2(0.f, 0.f); on[i] = make_float2(0.f, 0.f); }
  float2 sm2 = make_float2(0.f, 0.f);
  float amax = 0.f;

  const float* xb = a.audio + (long long)b * C * a.N;
  const bool vec_ok = ((a.N & 3) == 0) && ((hop & 3) == 0) && ((pad & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.audio) & 15) == 0);

  // software pipeline: tile i+1 streams into the other audio buffer while tile i
  // is transformed and accumulated
  unsigned par = 0;                 // mbarrier phase parity per buffer (bit b)
  bool async_cur = false;
  if (t_begin < t_end)
    async_cur = stage_tile_begin<C, TT>(sm, 0, xb, a.N, nb, t_begin, imin(TT, t_end - t_begin), hop,
                                        pad, vec_ok);
  // mask rows of this thread's bin: one frame per step of `mstride`.  They are
  // fetched with 4-byte cp.async straight into shared memory at the top of the
  // tile and first touched after the FFT phase, so their latency costs neither
  // registers nor issue slots
  const bool has_mn = a.mask_n != nullptr;
  const bool clip = (a.flags & SETK_F_CLIP_MASK) != 0;
  const long long mstride = (a.flags & SETK_F_MASK_FT) ? 1 : F;
  const long long mbase = (a.flags & SETK_F_MASK_FT) ? ((long long)b * F + (cov_thread ? bin : 0)) * a.T
                                                     : (long long)b * a.T * F + (cov_thread ? bin : 0);
  const float* mps = a.mask_s + mbase + (long long)t_begin * mstride;
  const float* mpn = has_mn ? a.mask_n + mbase + (long long)t_begin * mstride : nullptr;
  float* s_mask = sm.end();                        // [TT][2][MPITCH]
  constexpr int MPITCH = 260;
  int buf = 0;
  for (int t0 = t_begin; t0 < t_end; t0 += TT, buf ^= 1) {
    const int nt = imin(TT, t_end - t0);
    __syncthreads();   // tile i-1 fully consumed: sm.z and audio[buf^1] are free
    bool async_next = false;
    if (t0 + TT < t_end)
      async_next = stage_tile_begin<C, TT>(sm, buf ^ 1, xb, a.N, nb, t0 + TT,
                                           imin(TT, t_end - t0 - TT), hop, pad, vec_ok);
    if (cov_thread) {
#pragma unroll
      for (int j = 0; j < TT; ++j) {
        if (j < nt) {
          cp_async_f32(s_mask + (2 * j) * MPITCH + bin, mps + j * mstride);
          if (has_mn) cp_async_f32(s_mask + (2 * j + 1) * MPITCH + bin, mpn + j * mstride);
        }
      }
      mps += TT * mstride;
      if (has_mn) mpn += TT * mstride;
    }
    if (async_cur) {
      mbar_wait(&sm.bar[buf], (par >> buf) & 1u);
      par ^= 1u << buf;
    }
    if (warp < 8) fft_tile<C, TT>(sm, buf, nt, hop, w1, amax);
    async_cur = async_next;
    cp_async_wait_all();
    __syncthreads();
    // ---- covariance: thread per bin ----
    if (cov_thread) {
#pragma unroll
      for (int j = 0; j < TT; ++j) {
        if (j < nt) {
          float2 x[C];
#pragma unroll
          for (int c = 0; c < C; ++c) {
            const float2* z = sm.z + (j * C + c) * SETK_ZSLOT;
            x[c] = split_bin(z[zk], z[zn], tw);
          }
          if (bin == 0 || bin == kM) {
#pragma unroll
            for (int c = 0; c < C; ++c) x[c].y = 0.f;               // DC / Nyquist are real
          }
          const float m_raw = s_mask[(2 * j) * MPITCH + bin];
          const float m_s = clip ? fminf(m_raw, 1.0f) : m_raw;
          const float m_n = has_mn ? s_mask[(2 * j + 1) * MPITCH + bin] : 1.0f - m_s;
          const float2 msn = make_float2(m_s, m_n);
          const float2 mss = make_float2(m_s, m_s), mnn = make_float2(m_n, m_n);
          sm2 = f2add(sm2, msn);
          int o = 0;
#pragma unroll
          for (int i = 0; i < C; ++i) {
            const float pii = x[i].x * x[i].x + x[i].y * x[i].y;
            dg[i] = f2fma(msn, make_float2(pii, pii), dg[i]);
#pragma unroll
            for (int k = i + 1; k < C; ++k) {
              const float2 pr = cmul_conj(x[i], x[k]);        // x_i conj(x_k): two packed instructions
              os[o] = f2fma(pr, mss, os[o]);
              on[o] = f2fma(pr, mnn, on[o]);
              ++o;
            }
          }
        }
      }
    }
  }

  // center=False leaves a tail no frame covers; max|x| must still see it
  if (a.maxabs_bits && chunk == a.n_chunks - 1) {
    const int covered = (Tb > 0 ? (Tb - 1) * hop + kNfft - 2 * pad : 0);
    for (int c = 0; c < C; ++c)
      for (int i = imax(covered, 0) + tid; i < nb; i += blockDim.x)
        amax = fmaxf(amax, fabsf(xb[(long long)c * a.N + i]));
  }

  // ---- write partial sums ----
  if (cov_thread) {
    float* pp = a.partials + (((long long)b * a.n_chunks + chunk) * (2 * NACC + 2)) * F + bin;
#pragma unroll
    for (int i = 0; i < C; ++i) { pp[(long long)i * F] = dg[i].x; pp[(long long)(NACC + i) * F] = dg[i].y; }
#pragma unroll
    for (int i = 0; i < NOFF; ++i) {
      pp[(long long)(C + 2 * i) * F] = os[i].x;        pp[(long long)(C + 2 * i + 1) * F] = os[i].y;
      pp[(long long)(NACC + C + 2 * i) * F] = on[i].x; pp[(long long)(NACC + C + 2 * i + 1) * F] = on[i].y;
    }
    pp[(long long)(2 * NACC) * F] = sm2.x;
    pp[(long long)(2 * NACC + 1) * F] = sm2.y;
  }
  if (a.maxabs_bits) {
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if (lane == 0 && amax > 0.f) atomicMax(a.maxabs_bits + b, __float_as_uint(amax));
  }
}

// Deterministic reduction over chunks + normalisation + Hermitian fill.
// One thread per (b, f).
template <int C>
__global__ void cov_finalize_kernel(const float* __restrict__ partials, int B, int F, int n_chunks,
                                    float2* __restrict__ Rs, float2* __restrict__ Rn) {
  constexpr int NACC = CovAcc<C>::NACC;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * F) return;
  const int b = (int)(idx / F), f = (int)(idx % F);
  float acc[2 * NACC + 2];
#pragma unroll
  for (int i = 0; i < 2 * NACC + 2; ++i) acc[i] = 0.f;
  for (int ch = 0; ch < n_chunks; ++ch) {
    const float* pp = partials + (((long long)b * n_chunks + ch) * (2 * NACC + 2)) * F + f;
#pragma unroll
    for (int i = 0; i < 2 * NACC + 2; ++i) acc[i] += pp[(long long)i * F];
  }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const float* A = acc + which * NACC;
    const float inv = 1.0f / fmaxf(acc[2 * NACC + which], 1e-6f);
    float2* R = (which == 0 ? Rs : Rn) + idx * (C * C);
    int o = C;
#pragma unroll
    for (int i = 0; i < C; ++i) {
      R[i * C + i] = make_float2(A[i] * inv, 0.f);
#pragma unroll
      for (int k = i + 1; k < C; ++k) {
        const float re = A[o] * inv, im = A[o + 1] * inv;
        R[i * C + k] = make_float2(re, im);
        R[k * C + i] = make_float2(re, -im);
        o += 2;
      }
    }
  }
}

cudaError_t run_bits_to_float(const unsigned* bits, int n, float* out, void* stream);

template <int C, int TT>
static size_t stft_cov_smem_bytes(int hop) {
  return sizeof(float) * (TileSmem<C, TT>::floats(hop) + (size_t)TT * 2 * 260);
}

template <int C, int TT>
static cudaError_t run_stft_cov_t(setk_plan* pl, StftCovArgs a, int B, float2* Rs, float2* Rn,
                                  float* maxabs, void* stream) {
  const size_t smem = stft_cov_smem_bytes<C, TT>(a.g.hop);
  cudaError_t e = cudaFuncSetAttribute(stft_cov_kernel<C, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  e = launch(stft_cov_kernel<C, TT>, dim3(a.n_chunks, B), dim3(288), smem, stream, false, a);
  if (e != cudaSuccess) return e;
  const long long n = (long long)B * a.g.F;
  e = launch(cov_finalize_kernel<C>, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, stream, true,
             (const float*)a.partials, B, a.g.F, a.n_chunks, Rs, Rn);
  if (e != cudaSuccess) return e;
  if (maxabs) e = run_bits_to_float(a.maxabs_bits, B, maxabs, stream);
  return e;
}

// Does a fused instantiation exist for this geometry?
bool stft_cov_fused_supported(const Geometry& g) {
  if (g.n_fft != 512) return false;
  if (g.C < 1 || g.C > 4) return false;
  if (g.hop < 2 || g.hop > 512 || (g.hop & 1)) return false;   // float2 frame loads
  return true;
}

// partial-sum workspace: floats per (utterance, chunk)
size_t stft_cov_partial_floats(const Geometry& g) { return (size_t)(2 * g.C * g.C + 2) * g.F; }

int stft_cov_pick_chunks(const setk_plan* pl, int B, int T) {
  // aim for >= 2 waves of (2 CTAs / SM) with at most 16 chunks per utterance
  const int slots = 2 * pl->sm_count;
  int chunks = (2 * slots + B - 1) / B;
  if (chunks < 1) chunks = 1;
  if (chunks > 16) chunks = 16;
  const int max_chunks = (T + 3) / 4;
  if (chunks > max_chunks) chunks = max_chunks;
  return chunks < 1 ? 1 : chunks;
}

cudaError_t run_stft_cov_fused(setk_plan* pl, const float* audio, const int* n_samples, int B, int N,
                               int T, const float* mask_s, const float* mask_n, unsigned flags,
                               int n_chunks, float* partials, unsigned* maxabs_bits, float2* Rs,
                               float2* Rn, float* maxabs, void* stream) {
  constexpr int TT = 4;
  StftCovArgs a;
  a.g = pl->geo;
  a.audio = audio; a.n_samples = n_samples; a.N = N;
  a.mask_s = mask_s; a.mask_n = mask_n; a.flags = flags;
  a.T = T;
  a.n_chunks = n_chunks;
  int fpc = (T + n_chunks - 1) / n_chunks;
  a.frames_per_chunk = ((fpc + TT - 1) / TT) * TT;
  a.window = pl->d_window;
  a.partials = partials;
  a.maxabs_bits = maxabs_bits;
  switch (pl->geo.C) {
    case 1: return run_stft_cov_t<1, TT>(pl, a, B, Rs, Rn, maxabs, stream);
    case 2: return run_stft_cov_t<2, TT>(pl, a, B, Rs, Rn, maxabs, stream);
    case 3: return run_stft_cov_t<3, TT>(pl, a, B, Rs, Rn, maxabs, stream);
    case 4: return run_stft_cov_t<4, TT>(pl, a, B, Rs, Rn, maxabs, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace setk
