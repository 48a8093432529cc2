/*
 * setk_b200.h -- C-ABI of libsetk_b200.so: the B200-native (sm_100a) mask-based
 * adaptive-beamformer hot path of funcwj/setk.
 *
 *   multichannel STFT -> mask-weighted spatial covariance -> MVDR / MPDR / GEV /
 *   PMWF weight solve (+BAN, rank-1) -> beamform apply -> iSTFT (+peak norm)
 *
 * Every entry point below replaces one reference interface; the citation is
 * the reference file:line (paths relative to the reference tree).  The Python
 * mirror of the reference's operator API (setk_b200/libs/{utils,stft,
 * beamformer}.py) binds these with ctypes; INTEGRATION.md shows the stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers unless the name ends in _host;
 *   - every call is asynchronous on the caller's `stream` (a cudaStream_t passed
 *     as void*), re-entrant, and keeps no global state besides the last error
 *     string (thread local);
 *   - arrays are dense, row-major, the reference's axis order:
 *       audio  f32 [B][C][N]          (read_wav: C x N, utils.py:65-92)
 *       mask   f32 [B][T][F]          (T x F, beamformer.py:89)
 *       stft   c64 [B][C][F][T]       (N x F x T, data_handler.py:502-503)
 *       R      c64|c128 [B][F][C][C]  (F x N x N, beamformer.py:93)
 *       weight c64|c128 [B][F][C]     (F x N,     beamformer.py:18)
 *       enh    c64 [B][F][T]          (F x T,     beamformer.py:226)
 *       wave   f32 [B][N_out]
 *     complex = interleaved (re, im);
 *   - ragged batches: `n_samples` (i32[B], device) gives each utterance's true
 *     length (<= N); NULL means every utterance has N samples.  Frames past an
 *     utterance's own frame count are skipped (covariance) / written as zero;
 *   - return value: 0 ok, < 0 bad argument (SETK_E*), > 0 a cudaError_t.
 *     Numerical failures (singular Rn, non-PD Rn for GEV, ...) never fail the
 *     call: they set bits in the per-utterance `status` word, which the Python
 *     layer maps to numpy.linalg.LinAlgError per key like the reference CLI's
 *     per-utterance try/except (apply_adaptive_beamformer.py:160-172).
 */
#ifndef SETK_B200_H_
#define SETK_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SETK_VERSION 100 /* 0.1.0 */

#define SETK_MAX_CHANNELS 16

/* ---- error codes (negative) ---- */
#define SETK_OK 0
#define SETK_EINVAL -1      /* bad argument / unsupported configuration        */
#define SETK_ENOMEM -2      /* host allocation failed                           */
#define SETK_ESHAPE -3      /* shape inconsistent with the plan                 */
#define SETK_EUNSUPPORTED -4

/* ---- per-utterance status bits written by setk_weights ---- */
#define SETK_ST_SINGULAR 1u     /* exactly singular pivot in a linear solve (np.linalg.solve -> LinAlgError) */
#define SETK_ST_NOT_PD 2u       /* Rn not positive definite in the GEV reduction (scipy eigh(a,b) -> LinAlgError) */
#define SETK_ST_NO_CONVERGE 4u  /* Jacobi eigen-iteration hit its sweep limit  */
#define SETK_ST_NONFINITE 8u    /* NaN/Inf in the result                       */
#define SETK_ST_BAD_REF 16u     /* PMWF reference channel >= C (RuntimeError, beamformer.py:656-658) */
#define SETK_ST_REGULARIZED 32u /* WARNING only: Rn was numerically not PD in a GEV reduction and its
                                   diagonal was loaded (the reference falls back to a non-Hermitian eig
                                   there, beamformer.py:55-58) */

/* ---- beamformer kinds: apply_adaptive_beamformer.py:22,92-110 ---- */
#define SETK_BF_MVDR 0          /* beamformer.py:527-539 */
#define SETK_BF_MPDR 1          /* beamformer.py:555-573 (Ry in place of Rn)   */
#define SETK_BF_MPDR_WHITEN 2   /* beamformer.py:566-568 */
#define SETK_BF_GEVD 3          /* beamformer.py:674-682 */
#define SETK_BF_PMWF 4          /* beamformer.py:632-659 */
#define SETK_BF_PEVD 5          /* bare solve_pevd(Rs) / solve_pevd(Rs,Rn) when Rn != NULL (beamformer.py:31-63) */

/* ---- rank-1 approximation of Rs for PMWF: beamformer.py:66-84,641-645 ---- */
#define SETK_RANK1_NONE 0
#define SETK_RANK1_EIG 1
#define SETK_RANK1_GEV 2

/* ---- flags for setk_stft_cov / setk_cov ---- */
#define SETK_F_CLIP_MASK 1u     /* mask <- min(mask, 1): apply_adaptive_beamformer.py:142 */
#define SETK_F_MASK_FT 2u       /* masks are stored [B][F][T] and transposed on the fly (ibid. 149-152) */

/* ---- element types of R / weight buffers ---- */
#define SETK_C64 0              /* complex<float>  */
#define SETK_C128 1             /* complex<double> */

typedef struct setk_plan setk_plan_t;

/*
 * STFT geometry of one plan.  Mirrors StftParser (libs/opts.py:21-49) and the
 * keyword arguments of forward_stft / inverse_stft (libs/utils.py:96-105,
 * 142-150).  n_fft = nextpow2(frame_len) when round_power_of_two, else
 * frame_len (utils.py:114); only power-of-two n_fft in [32, 4096] is supported
 * (SETK_EUNSUPPORTED otherwise).
 */
typedef struct setk_config {
  int32_t num_channels;      /* C, 1..SETK_MAX_CHANNELS                                   */
  int32_t frame_len;         /* window length (--frame-len)                               */
  int32_t n_fft;             /* FFT size >= frame_len                                     */
  int32_t frame_hop;         /* --frame-hop                                               */
  int32_t center;            /* --center: reflect-pad n_fft/2 both sides                  */
  int32_t max_batch;         /* largest B any call will pass (workspace sizing)           */
  int32_t max_samples;       /* largest N any call will pass                              */
  int32_t reserved;
  const double* window_host; /* HOST pointer, frame_len analysis-window values
                                (scipy.signal.get_window(window, frame_len, fftbins=True),
                                or hann**0.5 for "sqrthann": utils.py:116-117)            */
} setk_config_t;

/* -------------------------------------------------------------------------- */
int setk_version(void);
/* thread-local description of the last non-zero return */
const char* setk_last_error_string(void);

/* Plan: window tables, synthesis-normalisation tables, partial-sum and STFT
 * workspace.  One per (C, frame_len, n_fft, hop, center, max_batch, max_samples).
 * Replaces the per-call window construction of librosa.stft/istft reached from
 * utils.py:123-128,159-164 and stft.h:75-154 (ShortTimeFTComputer). */
int setk_plan_create(const setk_config_t* cfg, setk_plan_t** plan_out);
int setk_plan_destroy(setk_plan_t* plan);

/* Integer bookkeeping, bit-exact with librosa 0.8.1 and
 * stft.h:143-153 (NumFrames / NumSamples):
 *   T     = 1 + (N + 2*(n_fft/2)*center - n_fft) / hop      (or -1 if too short)
 *   N_out = hop*(T-1) + n_fft - 2*(n_fft/2)*center */
int setk_num_frames(const setk_plan_t* plan, int32_t n_samples);
int setk_istft_length(const setk_plan_t* plan, int32_t num_frames);
int setk_num_bins(const setk_plan_t* plan);

/* forward_stft x C + np.stack: utils.py:96-138, data_handler.py:492-503;
 * C++ twin ShortTimeFTComputer::Compute (stft.cc:28-66,108).
 * stft_out c64 [B][C][F][T], T = setk_num_frames(plan, N). */
int setk_stft(setk_plan_t* plan, const float* audio, const int32_t* n_samples,
              int32_t B, int32_t N, void* stft_out, void* stream);

/* THE METRIC KERNEL.  Fused multichannel STFT + mask-weighted spatial
 * covariance: data_handler.py:502-503 -> beamformer.py:279-281 -> 87-103
 * (SupervisedBeamformer.run computing Rn then Rs), C++ twin EstimatePsd
 * (beamformer.cc:91-120).  The STFT is never written to HBM.
 *   mask_s  f32 [B][T][F]  target mask
 *   mask_n  f32 [B][T][F] or NULL -> 1 - mask_s (beamformer.py:279)
 *   Rs, Rn  c64 [B][F][C][C]: sum_t m x x^H / max(sum_t m, 1e-6)
 *   maxabs  f32 [B] or NULL: max |sample| over all channels
 *           (SpectrogramReader.maxabs, data_handler.py:398-400) -- the `norm`
 *           later handed to inverse_stft (apply_adaptive_beamformer.py:133,176) */
int setk_stft_cov(setk_plan_t* plan, const float* audio, const int32_t* n_samples,
                  int32_t B, int32_t N, const float* mask_s, const float* mask_n,
                  uint32_t flags, void* Rs, void* Rn, float* maxabs, void* stream);

/* compute_covar on an explicit STFT: beamformer.py:87-103.  Plan-free.
 *   stft c64 [B][C][F][T], mask f32 [B][T][F], R c64 [B][F][C][C] */
int setk_cov(const void* stft, const float* mask, uint32_t flags, int32_t B,
             int32_t C, int32_t F, int32_t T, void* R, void* stream);

/* Per-bin weight solve, fp64 internally, one thread-resident problem per
 * (utterance, bin).  Replaces solve_pevd / do_ban / rank1_constraint /
 * {Mvdr,Mpdr,Gevd,Pmwf}Beamformer.weight (beamformer.py:14-84,527-682) and the
 * C++ twins EstimateSteerVector / ComputeMvdrBeamWeights /
 * ComputeGevdBeamWeights (beamformer.cc:128-204).
 *   kind         SETK_BF_*
 *   beta         PMWF beta (0: pmwf-0, 1: pmwf-1)
 *   ref_channel  PMWF reference channel, < 0: pick argmax of the estimated
 *                output SNR (beamformer.py:620-630,650-653)
 *   rank1        SETK_RANK1_* (PMWF only)
 *   ban          != 0: blind analytical normalisation with Rn (beamformer.py:14-28)
 *   Rs, Rn, Ry   [B][F][C][C] of r_dtype (Ry only for MPDR kinds; Rn may be
 *                NULL for SETK_BF_PEVD and for SETK_BF_MPDR without ban)
 *   w            [B][F][C] of w_dtype
 *   status       u32 [B], OR-ed SETK_ST_* bits (caller zeroes it)
 *   ref_used     i32 [B] or NULL: PMWF reference channel actually used
 * Eigenvector convention (the reference's is LAPACK-defined up to a sign, see
 * SURVEY.md finding 4): principal eigenvector has unit 2-norm and component 0
 * real and >= 0; GEV: Rn = L L^H, y principal of L^-1 Rs L^-H with the same
 * convention, w = L^-H y (so w^H Rn w = 1, scipy.linalg.eigh(a, b)). */
int setk_weights(int32_t kind, double beta, int32_t ref_channel, int32_t rank1,
                 int32_t ban, const void* Rs, const void* Rn, const void* Ry,
                 int32_t r_dtype, int32_t B, int32_t F, int32_t C, void* w,
                 int32_t w_dtype, uint32_t* status, int32_t* ref_used, void* stream);

/* do_ban(weight, Rn): beamformer.py:14-28.  w_in / w_out [B][F][C], Rn [B][F][C][C],
 * all of `dtype` (SETK_C64 / SETK_C128). */
int setk_ban(const void* w_in, const void* Rn, int32_t dtype, int32_t B, int32_t F, int32_t C,
             void* w_out, void* stream);

/* rank1_constraint(Rs, Rn=None): beamformer.py:66-84.  Rn == NULL: principal
 * eigenvector of Rs; else Rn * (principal generalised eigenvector of (Rs, Rn)).
 * R1_out [B][F][C][C] = v v^H * tr(Rs) / max(tr(v v^H), eps32).  status u32[B] or NULL. */
int setk_rank1(const void* Rs, const void* Rn, int32_t dtype, int32_t B, int32_t F, int32_t C,
               void* R1_out, uint32_t* status, void* stream);

/* Beamformer.beamform on an explicit STFT: beamformer.py:220-234, C++ twin
 * Beamform (beamformer.cc:215-230).  enh[b][f][t] = sum_c conj(w[b][f][c]) x[b][c][f][t]
 *   post_mask f32 [B][T][F] or NULL: enh *= mask^T (apply_adaptive_beamformer.py:174-175) */
int setk_apply(const void* stft, const void* w, int32_t w_dtype, const float* post_mask,
               int32_t B, int32_t C, int32_t F, int32_t T, void* enh, void* stream);

/* inverse_stft: utils.py:142-173 (librosa.istft + peak normalisation), C++ twin
 * InverseShortTimeFT (stft.cc:154-198).
 *   enh     c64 [B][F][T]
 *   n_out   output samples per utterance: setk_istft_length(plan, T) when the
 *           reference passes nsamps=None, else nsamps (fix_length semantics)
 *   norm    f32 [B] or NULL: if given, wave <- wave * norm / (max|wave| + eps32)
 *           (utils.py:166-168)
 *   wave    f32 [B][n_out] */
int setk_istft(setk_plan_t* plan, const void* enh, int32_t B, int32_t T, int32_t n_out,
               const float* norm, float* wave, void* stream);

/* Fused beamform-apply + iSTFT from the multichannel AUDIO (the STFT is
 * recomputed on chip, never read from HBM): beamformer.py:283 ->
 * apply_adaptive_beamformer.py:174-176 -> utils.py:159-168.
 *   w        [B][F][C] of w_dtype
 *   post_mask, norm, wave, n_out as above */
int setk_apply_istft(setk_plan_t* plan, const float* audio, const int32_t* n_samples,
                     int32_t B, int32_t N, const void* w, int32_t w_dtype,
                     const float* post_mask, int32_t n_out, const float* norm,
                     float* wave, void* stream);

/* The same, through to the wav writer: the pass that applies inverse_stft's `norm`
 * rescale (utils.py:166-168) also performs WaveWriter.write -> write_wav's float ->
 * PCM-16 conversion (data_handler.py:600-605, utils.py:45-62; soundfile's default
 * subtype = floor(y * 32768) clipped, SURVEY.md finding 3), so the float wave is
 * never written out: bit-identical to setk_apply_istft + setk_float_to_pcm16.
 *   norm  f32 [B] or NULL (no rescale)      pcm  i16 [B][n_out] */
int setk_apply_istft_pcm16(setk_plan_t* plan, const float* audio, const int32_t* n_samples,
                           int32_t B, int32_t N, const void* w, int32_t w_dtype,
                           const float* post_mask, int32_t n_out, const float* norm,
                           int16_t* pcm, void* stream);

/*
 * CGMM time-frequency mask estimation from audio: the tile STFT, then
 * CgmmTrainer(stft, num_classes, gamma=init, update_alpha=...).train(num_iters)
 * (scripts/sptk/libs/cluster.py:396-465 with 94-130, 187-287) for every utterance
 * of the batch, and the CLI's transpose (scripts/sptk/estimate_cgmm_masks.py:36-60).
 * fp64 arithmetic after the complex64 STFT, like the reference.
 *
 *   num_classes  2..4.  init_gamma == NULL needs num_classes == 2 (the reference's
 *                deterministic start R_0 = sum_t y y^H / T, R_1 = I; for more classes
 *                it draws numpy random posteriors -- pass them in).
 *   init_gamma   f32 [B][K][T][F] starting posteriors or NULL
 *   update_alpha nonzero: the priors follow mean_t gamma (cluster.py:251-252)
 *   masks        f32 [B][K][T][F]  (the reference writes masks[0] when K == 2)
 *   status       u32 [B] or NULL, OR-ed SETK_ST_NO_CONVERGE (caller zeroes it)
 * Needs n_fft 512 or 1024 (the tile STFT); SETK_EUNSUPPORTED otherwise.
 */
int setk_cgmm_masks(setk_plan_t* plan, const float* audio, const int32_t* n_samples, int32_t B, int32_t N,
                    int32_t num_classes, int32_t num_iters, const float* init_gamma, int32_t update_alpha,
                    float* masks, uint32_t* status, void* stream);

/*
 * The same estimator for a caller that brings its own STFT -- the reference's
 * CgmmTrainer(obs, num_classes, gamma, update_alpha).train(num_iters) signature
 * (cluster.py:401-466).  Plan-free; scratch is stream-ordered (cudaMallocAsync).
 *   stft   c64 [B][C][F][T]  (forward_stft with transpose=False, stacked over channels)
 *   other arguments as for setk_cgmm_masks; any F, T >= 1.
 */
int setk_cgmm_stft(const void* stft, int32_t B, int32_t C, int32_t F, int32_t T, int32_t num_classes,
                   int32_t num_iters, const float* init_gamma, int32_t update_alpha, float* masks,
                   uint32_t* status, void* stream);

/*
 * WPE dereverberation of a multichannel STFT: libs/wpe.py wpe(reverb, taps, delay,
 * context, num_iters) (scripts/sptk/libs/wpe.py:82-110 with 14-79; driven by
 * apply_wpe.py:36-52) for every utterance of the batch.  Plan-free; scratch is
 * stream-ordered.  fp64 arithmetic (the reference computes in complex64 for
 * complex64 input).
 *   stft   c64 [B][C][F][T]   (the reference's F x N x T per utterance, channel-major here)
 *   out    c64 [B][C][F][T]   dereverberated
 *   status u32 [B] or NULL, SETK_ST_SINGULAR when solve(R, r) hits an exactly singular
 *          matrix (numpy.linalg.LinAlgError in the reference, caught per utterance by its CLI)
 * channels x taps <= 128 and the bin's time series must fit shared memory
 * (SETK_EUNSUPPORTED otherwise; 8 ch x 10 taps x 10 s at hop 256 does).
 */
int setk_wpe_stft(const void* stft, int32_t B, int32_t C, int32_t F, int32_t T, int32_t taps, int32_t delay,
                  int32_t context, int32_t num_iters, void* out, uint32_t* status, void* stream);

/* One WPE step with a caller-supplied variance: the dereverberation stage of facted_wpd
 * (libs/wpe.py:113-177; wpe_step 58-77 with lambda from the previous WPD output, 150-153).
 *   stft        c64 [B][C][F][T]
 *   lambda_enh  c64 [B][F][T] or NULL: lambda = max(|lambda_enh|^2, eps32); NULL: compute_lambda(stft, context)
 *               (then the call equals setk_wpe_stft with num_iters = 1)
 *   out         c64 [B][C][F][T]   dereverberated observations
 *   inv_lambda  f32 [B][T][F] or NULL: 1 / lambda, the frame weights of Rd (wpe.py:165) in the layout
 *               setk_cov takes as a mask
 *   status      u32 [B] or NULL (SETK_ST_SINGULAR like setk_wpe_stft) */
int setk_wpe_step(const void* stft, const void* lambda_enh, int32_t B, int32_t C, int32_t F, int32_t T,
                  int32_t taps, int32_t delay, int32_t context, void* out, float* inv_lambda,
                  uint32_t* status, void* stream);

/* ---- spatial features on explicit STFTs: scripts/sptk/libs/spatial.py ---- */

/* ipd(si, sj, cos, sin): spatial.py:163-181.  si, sj c64 [rows][F] (rows = B*T);
 *   mode 0: np.mod(angle(si) - angle(sj) + pi, 2 pi) - pi   -> out f32 [rows][F]
 *   mode 1: cos(angle(si) - angle(sj))                      -> out f32 [rows][F]
 *   mode 2: [cos | sin]                                     -> out f32 [rows][2F] */
int setk_ipd(const void* si, const void* sj, int64_t rows, int32_t F, int32_t mode, float* out,
             void* stream);

/* directional_feats(spectrogram, steer_vector, df_pair): spatial.py:184-208.
 *   stft   c64  [B][M][F][T]
 *   steer  c128 [M][F] (steer_batched == 0) or [B][M][F]
 *   pairs  i32  [n_pairs][2] or NULL (all i < j)
 *   out    f64  [B][T][F] = mean over pairs of cos((arg x_i - arg x_j) - (arg d_i - arg d_j)) */
int setk_directional_feats(const void* stft, const void* steer, int32_t steer_batched,
                           const int32_t* pairs, int32_t n_pairs, int32_t B, int32_t M, int32_t F,
                           int32_t T, double* out, void* stream);

/* gcc_phat_linear / gcc_phat_diag: spatial.py:37-92 (the caller supplies the TDOA grid that
 * linear_tdoa_grid, 11-34, or the circular-array formula, 78-80, produce).
 *   si, sj c64 [T][F];  omega f64 [F] (rad/s);  tau f64 [D] (s)
 *   spectrum = Re(exp(j (angle si - angle sj)) @ exp(-j outer(omega, tau)))      [T][D]
 *   normalize:   spectrum /= max(max |spectrum|, eps32);  apply_floor: max(spectrum, 0)
 *   accumulate != 0: out += spectrum (srp_phat_linear, 95-123, sums the pairs) else out = spectrum
 *   work   f64 [setk_gcc_phat_work_doubles(T, F, D)] scratch;  out f64 [T][D] */
int64_t setk_gcc_phat_work_doubles(int32_t T, int32_t F, int32_t D);
int setk_gcc_phat(const void* si, const void* sj, int32_t T, int32_t F, const double* omega,
                  const double* tau, int32_t D, int32_t normalize, int32_t apply_floor,
                  int32_t accumulate, double* work, double* out, void* stream);

/* msc(spectrogram, context, normalize): spatial.py:126-160, as written (the diagonal terms
 * enter as their grand total, np.sum without an axis at :153).
 *   spec c64 [N][T][F];  work f64 [setk_msc_work_doubles(T, F)];  out f64 [T][F] */
int64_t setk_msc_work_doubles(int32_t T, int32_t F);
int setk_msc(const void* spec, int32_t N, int32_t T, int32_t F, int32_t context, int32_t normalize,
             double* work, double* out, void* stream);

/* floor(y * 32768) clipped to int16: the PCM_16 conversion of
 * WaveWriter.write -> write_wav -> soundfile (data_handler.py:600-605,
 * utils.py:45-62; SURVEY.md finding 3).  wave f32 [n], pcm i16 [n]. */
int setk_float_to_pcm16(const float* wave, int64_t n, int16_t* pcm, void* stream);
/* int16 / 32768 -> float32: read_wav (utils.py:80-92). */
int setk_pcm16_to_float(const int16_t* pcm, int64_t n, float* wave, void* stream);

/* Kaldi CompressedMatrix masks, format "CM" (per-column percentile headers + one byte per element),
 * expanded on the device: kaldi_io.py:248-281 uncompress() as reached by the mask ScriptReader of
 * apply_adaptive_beamformer.py:139-142 (--mask-format kaldi), same float32 operations, bit-identical.
 *   blobs  u8  [B][slot_bytes]  one matrix per slot exactly as it lies in the archive behind the "CM "
 *                               token: {f32 min, f32 range, i32 rows, i32 cols}, cols x 4 u16,
 *                               cols x rows u8 (column major); slot_bytes a multiple of 16
 *   out    f32 [B][T][F]        rows t < rows_b decoded, the rest zero (ragged batches)
 *   status i32 [B]              set to 1 where a header is inconsistent (cols != F, rows > T, the
 *                               matrix does not fit its slot); that matrix decodes to zeros */
int setk_cm_masks(const uint8_t* blobs, int64_t slot_bytes, int32_t B, int32_t T, int32_t F, float* out,
                  int32_t* status, void* stream);

/* How many kernels of this library have been launched by this process
 * (for bench.py's "gpu_launches"). */
int64_t setk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SETK_B200_H_ */
