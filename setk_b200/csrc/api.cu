// api.cu -- the C-ABI of libsetk_b200.so (include/setk_b200.h): argument
// checking, plan / workspace management and dispatch to the kernels.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <vector>

namespace setk {

std::atomic<long long> g_launch_count{0};

// --- launchers implemented in the other translation units ---
cudaError_t run_stft_generic(const setk_plan*, const float*, const int*, int, int, int, float2*, void*);
cudaError_t run_cov_generic(const float2*, const float*, unsigned, int, int, int, int, float2*, void*);
cudaError_t run_apply_generic(const float2*, const void*, int, const float*, int, int, int, int, float2*, void*);
cudaError_t run_istft_generic(const setk_plan*, const float2*, int, int, int, int, const int*, float*, float*,
                              unsigned*, void*);
cudaError_t run_peak_scale(float*, int, int, const float*, const unsigned*, void*);
cudaError_t run_float_to_pcm16(const float*, long long, int16_t*, void*);
cudaError_t run_peak_scale_pcm16(const float*, int, int, const float*, const unsigned*, int16_t*, void*);
cudaError_t run_pcm16_to_float(const int16_t*, long long, float*, void*);
cudaError_t run_cm_masks(const unsigned char*, long long, int, int, int, float*, int*, void*);

cudaError_t run_ipd(const float2*, const float2*, long long, int, int, float*, void*);
cudaError_t run_dirfeat(const float2*, const double2*, int, const int*, int, int, int, int, int, double*,
                        void*);
size_t gcc_phat_work_doubles(int, int, int);
cudaError_t run_gcc_phat(const float2*, const float2*, int, int, const double*, const double*, int, int,
                         int, int, double*, double*, void*);
size_t msc_work_doubles(int, int);
cudaError_t run_msc(const float2*, int, int, int, int, int, double*, double*, void*);
bool stft_cov_fused_supported(const Geometry&);
size_t stft_cov_partial_floats(const Geometry&);
size_t stft_cov_partial_bytes(const setk_plan*, int, int);
int stft_cov_pick_chunks(const setk_plan*, int, int);
cudaError_t run_stft_cov_fused(setk_plan*, const float*, const int*, int, int, int, const float*,
                               const float*, unsigned, int*, float*, unsigned*, float2*, float2*, float*,
                               void*);
bool apply_istft_fused_supported(const Geometry&);
cudaError_t run_apply_istft_fused(setk_plan*, const float*, const int*, int, int, int, const void*, int,
                                  const float*, int, int*, float*, unsigned*, void*);

bool stft_spill_supported(const Geometry&);
size_t stft_spill_bytes(const Geometry&, int, int);
size_t apply_spill_bytes(const Geometry&, int, int);
cudaError_t run_spill_to_bcft(const setk_plan*, const float2*, int, int, float2*, void*);
struct CgCtx { Geometry geo; int sm_count; };
size_t cgmm_workspace_bytes(const CgCtx*, int, int, int, int);
cudaError_t run_cgmm(const CgCtx*, const float2*, int, double*, const int*, int, int, int, int, int, const float*,
                     int, float*, unsigned*, void*);
cudaError_t run_bcft_to_spill(const float2*, int, int, int, int, int, float2*, void*);
bool wpe_supported(int, int, int, int, int);
size_t wpe_workspace_bytes(int, int, int, int);
cudaError_t run_wpe(const float2*, int, int, int, int, int, int, int, int, int, double*, float2*, unsigned*,
                    const float2*, float*, void*);
cudaError_t run_apply_spill(setk_plan*, const float2*, const void*, int, const float*, int, int, float2*, void*);
cudaError_t run_istft_strided(const setk_plan*, const float2*, long long, long long, long long, int, int, int,
                              const int*, float*, float*, unsigned*, void*);
size_t cov_spill_partial_bytes(const Geometry&, int, int);
int cov_spill_chunks(const setk_plan*, int, int);
cudaError_t run_stft_spill(setk_plan*, const float*, const int*, int, int, int, int, float2*, unsigned*,
                           void*);
cudaError_t run_cov_spill(setk_plan*, const float2*, const float*, const float*, unsigned, const int*,
                          int, int, int, int, float*, float2*, float2*, void*);
cudaError_t run_bits_to_float(const unsigned*, int, float*, void*);

struct WeightsArgs;
cudaError_t weights_run(int kind, double beta, int ref_channel, int rank1, int ban, const void* Rs,
                        const void* Rn, const void* Ry, int r_dtype, int B, int F, int C, void* w,
                        int w_dtype, unsigned* status, int* ref_used, void* stream);
cudaError_t maxabs_generic(const float* audio, const int* n_samples, int B, int C, int N,
                           unsigned* bits, float* out, void* stream);
cudaError_t post_run(int mode, const void* A, const void* Rn, int dtype, int B, int F, int C, void* out,
                     unsigned* status, void* stream);

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_err, sizeof(g_err), "%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return (int)e;
}

// grow-only device workspace
template <class T>
static cudaError_t ensure(T** ptr, size_t* have, size_t want) {
  if (*have >= want && *ptr) return cudaSuccess;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *have = 0;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(ptr), want);
  if (e == cudaSuccess) *have = want;
  return e;
}

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace setk

using namespace setk;

extern "C" {

int setk_version(void) { return SETK_VERSION; }
#ifdef SETK_EMU
// present only in the CPU test tier's build (tests/emu): lets the Python layer
// tell the two apart so that the product library never sees host pointers
int setk_emulated(void) { return 1; }
#endif
const char* setk_last_error_string(void) { return g_err; }
int64_t setk_launch_count(void) { return (int64_t)g_launch_count.load(); }

int setk_plan_create(const setk_config_t* cfg, setk_plan_t** plan_out) {
  if (!cfg || !plan_out) return fail(SETK_EINVAL, "setk_plan_create: null argument");
  *plan_out = nullptr;
  if (cfg->num_channels < 1 || cfg->num_channels > SETK_MAX_CHANNELS)
    return fail(SETK_EINVAL, "num_channels %d outside [1, %d]", cfg->num_channels, SETK_MAX_CHANNELS);
  if (cfg->frame_len < 1 || cfg->frame_hop < 1)
    return fail(SETK_EINVAL, "frame_len %d / frame_hop %d must be positive", cfg->frame_len, cfg->frame_hop);
  if (cfg->n_fft < cfg->frame_len)
    return fail(SETK_EINVAL, "n_fft %d smaller than frame_len %d", cfg->n_fft, cfg->frame_len);
  // powers of two run the FFT kernels; any other EVEN size (--round-power-of-two false,
  // utils.py:115) the direct-DFT path of the generic kernels (librosa's istft itself derives
  // n_fft = 2 (F - 1), i.e. it only round-trips even sizes)
  if (cfg->n_fft < 32 || cfg->n_fft > 4096 || (cfg->n_fft & 1))
    return fail(SETK_EUNSUPPORTED, "n_fft %d: only even sizes in [32, 4096] are supported", cfg->n_fft);
  if (!cfg->window_host) return fail(SETK_EINVAL, "window_host is null");
  if (cfg->max_batch < 1 || cfg->max_samples < 1)
    return fail(SETK_EINVAL, "max_batch / max_samples must be positive");

  setk_plan* pl = static_cast<setk_plan*>(calloc(1, sizeof(setk_plan)));
  if (!pl) return fail(SETK_ENOMEM, "out of host memory");
  pl->cfg = *cfg;
  pl->cfg.window_host = nullptr;
  Geometry& g = pl->geo;
  g.C = cfg->num_channels;
  g.n_fft = cfg->n_fft;
  g.log2n = 0;
  while ((1 << g.log2n) < g.n_fft) ++g.log2n;
  if (!is_pow2(g.n_fft)) g.log2n = -1;     // marks the direct-DFT route
  g.hop = cfg->frame_hop;
  g.pad = cfg->center ? g.n_fft / 2 : 0;
  g.F = g.n_fft / 2 + 1;

  // librosa.util.pad_center: left pad (n_fft - frame_len) // 2
  std::vector<float> win(g.n_fft, 0.f), wsq(g.n_fft, 0.f);
  const int lpad = (g.n_fft - cfg->frame_len) / 2;
  for (int i = 0; i < cfg->frame_len; ++i) {
    const double w = cfg->window_host[i];
    win[lpad + i] = (float)w;
    wsq[lpad + i] = (float)(w * w);
  }
  {   // windows with a constant pair sum w[n] + w[n + n_fft/2] (Hann, Hamming, rectangular, ...)
    const int h = g.n_fft / 2;
    const double K = (double)win[0] + (double)win[h];
    bool ok = K > 1e-3;
    for (int i = 0; i < h && ok; ++i) ok = fabs((double)win[i] + (double)win[i + h] - K) <= 2e-7 * K;
    pl->win_pair_sum = ok ? (float)K : 0.f;
  }
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&pl->d_window), sizeof(float) * g.n_fft);
  if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&pl->d_wsq), sizeof(float) * g.n_fft);
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_window, win.data(), sizeof(float) * g.n_fft, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(pl->d_wsq, wsq.data(), sizeof(float) * g.n_fft, cudaMemcpyHostToDevice);
  int dev = 0;
  if (e == cudaSuccess) e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    setk_plan_destroy(pl);
    return cuda_fail(e, "setk_plan_create");
  }
  *plan_out = pl;
  return SETK_OK;
}

int setk_plan_destroy(setk_plan_t* pl) {
  if (!pl) return SETK_OK;
  if (pl->d_window) cudaFree(pl->d_window);
  if (pl->d_wsq) cudaFree(pl->d_wsq);
  if (pl->d_partials) cudaFree(pl->d_partials);
  if (pl->d_wave_ws) cudaFree(pl->d_wave_ws);
  if (pl->d_stft_ws) cudaFree(pl->d_stft_ws);
  if (pl->d_enh_ws) cudaFree(pl->d_enh_ws);
  if (pl->d_frames_ws) cudaFree(pl->d_frames_ws);
  if (pl->d_cgmm_ws) cudaFree(pl->d_cgmm_ws);
  if (pl->d_peak) cudaFree(pl->d_peak);
  if (pl->d_tile_prefix) cudaFree(pl->d_tile_prefix);
  free(pl);
  return SETK_OK;
}

int setk_num_frames(const setk_plan_t* pl, int32_t n_samples) {
  if (!pl) return fail(SETK_EINVAL, "null plan");
  const int T = frames_of(n_samples, pl->geo.n_fft, pl->geo.hop, pl->geo.pad);
  return T > 0 ? T : -1;
}

int setk_istft_length(const setk_plan_t* pl, int32_t num_frames) {
  if (!pl) return fail(SETK_EINVAL, "null plan");
  if (num_frames < 1) return -1;
  return pl->geo.hop * (num_frames - 1) + pl->geo.n_fft - 2 * pl->geo.pad;
}

int setk_num_bins(const setk_plan_t* pl) { return pl ? pl->geo.F : fail(SETK_EINVAL, "null plan"); }

static int check_batch(const setk_plan_t* pl, int B, int N, const char* who) {
  if (!pl) return fail(SETK_EINVAL, "%s: null plan", who);
  if (B < 1 || B > pl->cfg.max_batch) return fail(SETK_ESHAPE, "%s: batch %d outside [1, max_batch=%d]", who, B, pl->cfg.max_batch);
  if (N < 1 || N > pl->cfg.max_samples) return fail(SETK_ESHAPE, "%s: N=%d outside [1, max_samples=%d]", who, N, pl->cfg.max_samples);
  if (frames_of(N, pl->geo.n_fft, pl->geo.hop, pl->geo.pad) < 1)
    return fail(SETK_ESHAPE, "%s: N=%d too short for n_fft=%d", who, N, pl->geo.n_fft);
  return SETK_OK;
}

int setk_stft(setk_plan_t* pl, const float* audio, const int32_t* n_samples, int32_t B, int32_t N,
              void* stft_out, void* stream) {
  int rc = check_batch(pl, B, N, "setk_stft");
  if (rc) return rc;
  if (!audio || !stft_out) return fail(SETK_EINVAL, "setk_stft: null buffer");
  const Geometry& g = pl->geo;
  const int T = frames_of(N, g.n_fft, g.hop, g.pad);
  cudaError_t e;
  if (stft_spill_supported(g) && (long long)B * g.C <= 65535) {
    // tile STFT into the bin-major workspace, then one transposing copy
    e = ensure(&pl->d_stft_ws, &pl->stft_ws_bytes, stft_spill_bytes(g, B, T));
    if (e != cudaSuccess) return cuda_fail(e, "setk_stft(workspace)");
    const int groups = (g.C + 3) / 4;
    e = run_stft_spill(pl, audio, n_samples, B, N, T, stft_cov_pick_chunks(pl, B * groups, T), pl->d_stft_ws,
                       nullptr, stream);
    if (e == cudaSuccess)
      e = run_spill_to_bcft(pl, pl->d_stft_ws, B, T, static_cast<float2*>(stft_out), stream);
  } else {
    e = run_stft_generic(pl, audio, n_samples, B, N, T, static_cast<float2*>(stft_out), stream);
  }
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_stft");
}

int setk_cov(const void* stft, const float* mask, uint32_t flags, int32_t B, int32_t C, int32_t F,
             int32_t T, void* R, void* stream) {
  if (!stft || !mask || !R) return fail(SETK_EINVAL, "setk_cov: null buffer");
  if (B < 1 || F < 1 || T < 1 || C < 1 || C > SETK_MAX_CHANNELS)
    return fail(SETK_ESHAPE, "setk_cov: bad shape B=%d C=%d F=%d T=%d", B, C, F, T);
  if (B > 65535) return fail(SETK_ESHAPE, "setk_cov: batch %d > 65535", B);
  cudaError_t e = run_cov_generic(static_cast<const float2*>(stft), mask, flags, B, C, F, T,
                                  static_cast<float2*>(R), stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_cov");
}

int setk_stft_cov(setk_plan_t* pl, const float* audio, const int32_t* n_samples, int32_t B, int32_t N,
                  const float* mask_s, const float* mask_n, uint32_t flags, void* Rs, void* Rn,
                  float* maxabs, void* stream) {
  int rc = check_batch(pl, B, N, "setk_stft_cov");
  if (rc) return rc;
  if (!audio || !mask_s || !Rs || !Rn) return fail(SETK_EINVAL, "setk_stft_cov: null buffer");
  const Geometry& g = pl->geo;
  const int T = frames_of(N, g.n_fft, g.hop, g.pad);
  cudaError_t e = ensure(&pl->d_peak, &pl->peak_bytes, sizeof(unsigned) * 2 * (size_t)pl->cfg.max_batch);
  if (e != cudaSuccess) return cuda_fail(e, "setk_stft_cov(workspace)");
  unsigned* maxabs_bits = nullptr;
  if (maxabs) {
    maxabs_bits = pl->d_peak + pl->cfg.max_batch;   // second half: audio max|x|
    e = cudaMemsetAsync(maxabs_bits, 0, sizeof(unsigned) * B, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "setk_stft_cov(memset)");
  }
  if (stft_cov_fused_supported(g)) {
    e = ensure(&pl->d_partials, &pl->partials_bytes, stft_cov_partial_bytes(pl, B, T));
    if (e == cudaSuccess && n_samples)
      e = ensure(&pl->d_tile_prefix, &pl->tile_prefix_bytes, sizeof(int) * ((size_t)pl->cfg.max_batch + 1));
    if (e != cudaSuccess) return cuda_fail(e, "setk_stft_cov(workspace)");
    e = run_stft_cov_fused(pl, audio, n_samples, B, N, T, mask_s, mask_n, flags, pl->d_tile_prefix,
                           pl->d_partials, maxabs_bits, static_cast<float2*>(Rs),
                           static_cast<float2*>(Rn), maxabs, stream);
    return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_stft_cov");
  }
  if (stft_spill_supported(g)) {
    // many-channel route (C > 4 at n_fft = 512, any C at n_fft = 1024): fast STFT
    // into a bin-major workspace, then a streaming covariance kernel (stft_spill.cu)
    const int groups = (g.C + 3) / 4;
    const int chunks_a = stft_cov_pick_chunks(pl, B * groups, T);
    // >= 16 short fp32 runs per utterance, combined in double by the finalize kernel
    const int chunks_b = cov_spill_chunks(pl, B, T);
    e = ensure(&pl->d_stft_ws, &pl->stft_ws_bytes, stft_spill_bytes(g, B, T));
    if (e == cudaSuccess)
      e = ensure(&pl->d_partials, &pl->partials_bytes, cov_spill_partial_bytes(g, B, chunks_b));
    if (e != cudaSuccess) return cuda_fail(e, "setk_stft_cov(workspace)");
    e = run_stft_spill(pl, audio, n_samples, B, N, T, chunks_a, pl->d_stft_ws, maxabs_bits, stream);
    if (e == cudaSuccess)
      e = run_cov_spill(pl, pl->d_stft_ws, mask_s, mask_n, flags, n_samples, B, N, T, chunks_b,
                        pl->d_partials, static_cast<float2*>(Rs), static_cast<float2*>(Rn), stream);
    if (e == cudaSuccess && maxabs) e = run_bits_to_float(maxabs_bits, B, maxabs, stream);
    return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_stft_cov(spill)");
  }
  // shape-generic route: explicit STFT in workspace, then two covariance passes
  const size_t want = sizeof(float2) * (size_t)B * g.C * g.F * T;
  e = ensure(&pl->d_stft_ws, &pl->stft_ws_bytes, want);
  if (e != cudaSuccess) return cuda_fail(e, "setk_stft_cov(workspace)");
  e = run_stft_generic(pl, audio, n_samples, B, N, T, pl->d_stft_ws, stream);
  if (e == cudaSuccess)
    e = run_cov_generic(pl->d_stft_ws, mask_s, flags, B, g.C, g.F, T, static_cast<float2*>(Rs), stream);
  if (e == cudaSuccess) {
    if (mask_n)
      e = run_cov_generic(pl->d_stft_ws, mask_n, flags & ~SETK_F_CLIP_MASK, B, g.C, g.F, T,
                          static_cast<float2*>(Rn), stream);
    else
      e = run_cov_generic(pl->d_stft_ws, mask_s, flags | SETK_F_ONE_MINUS_INTERNAL, B, g.C, g.F, T,
                          static_cast<float2*>(Rn), stream);
  }
  if (e == cudaSuccess && maxabs)
    e = maxabs_generic(audio, n_samples, B, g.C, N, maxabs_bits, maxabs, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_stft_cov(generic)");
}

int setk_weights(int32_t kind, double beta, int32_t ref_channel, int32_t rank1, int32_t ban,
                 const void* Rs, const void* Rn, const void* Ry, int32_t r_dtype, int32_t B, int32_t F,
                 int32_t C, void* w, int32_t w_dtype, uint32_t* status, int32_t* ref_used, void* stream) {
  if (!Rs || !w || !status) return fail(SETK_EINVAL, "setk_weights: null buffer");
  if (B < 1 || F < 1 || C < 1 || C > SETK_MAX_CHANNELS)
    return fail(SETK_ESHAPE, "setk_weights: bad shape B=%d F=%d C=%d", B, F, C);
  if (kind < SETK_BF_MVDR || kind > SETK_BF_PEVD) return fail(SETK_EINVAL, "setk_weights: unknown kind %d", kind);
  if ((r_dtype != SETK_C64 && r_dtype != SETK_C128) || (w_dtype != SETK_C64 && w_dtype != SETK_C128))
    return fail(SETK_EINVAL, "setk_weights: bad dtype");
  const bool need_rn = kind == SETK_BF_MVDR || kind == SETK_BF_MPDR_WHITEN || kind == SETK_BF_GEVD ||
                       kind == SETK_BF_PMWF;
  if (need_rn && !Rn) return fail(SETK_EINVAL, "setk_weights: Rn required for kind %d", kind);
  if ((kind == SETK_BF_MPDR || kind == SETK_BF_MPDR_WHITEN) && !Ry)
    return fail(SETK_EINVAL, "setk_weights: Ry required for MPDR");
  if (ban && !Rn) return fail(SETK_EINVAL, "setk_weights: BAN needs Rn");
  if (rank1 < SETK_RANK1_NONE || rank1 > SETK_RANK1_GEV) return fail(SETK_EINVAL, "setk_weights: bad rank1 %d", rank1);
  cudaError_t e = weights_run(kind, beta, ref_channel, rank1, ban, Rs, Rn, Ry, r_dtype, B, F, C, w, w_dtype,
                              status, ref_used, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_weights");
}

int setk_ban(const void* w_in, const void* Rn, int32_t dtype, int32_t B, int32_t F, int32_t C,
             void* w_out, void* stream) {
  if (!w_in || !Rn || !w_out) return fail(SETK_EINVAL, "setk_ban: null buffer");
  if (B < 1 || F < 1 || C < 1 || C > SETK_MAX_CHANNELS) return fail(SETK_ESHAPE, "setk_ban: bad shape");
  if (dtype != SETK_C64 && dtype != SETK_C128) return fail(SETK_EINVAL, "setk_ban: bad dtype");
  cudaError_t e = post_run(0, w_in, Rn, dtype, B, F, C, w_out, nullptr, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_ban");
}

int setk_rank1(const void* Rs, const void* Rn, int32_t dtype, int32_t B, int32_t F, int32_t C,
               void* R1_out, uint32_t* status, void* stream) {
  if (!Rs || !R1_out) return fail(SETK_EINVAL, "setk_rank1: null buffer");
  if (B < 1 || F < 1 || C < 1 || C > SETK_MAX_CHANNELS) return fail(SETK_ESHAPE, "setk_rank1: bad shape");
  if (dtype != SETK_C64 && dtype != SETK_C128) return fail(SETK_EINVAL, "setk_rank1: bad dtype");
  cudaError_t e = post_run(1, Rs, Rn, dtype, B, F, C, R1_out, status, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_rank1");
}

int setk_apply(const void* stft, const void* w, int32_t w_dtype, const float* post_mask, int32_t B,
               int32_t C, int32_t F, int32_t T, void* enh, void* stream) {
  if (!stft || !w || !enh) return fail(SETK_EINVAL, "setk_apply: null buffer");
  if (B < 1 || F < 1 || T < 1 || C < 1 || C > SETK_MAX_CHANNELS)
    return fail(SETK_ESHAPE, "setk_apply: bad shape B=%d C=%d F=%d T=%d", B, C, F, T);
  cudaError_t e = run_apply_generic(static_cast<const float2*>(stft), w, w_dtype, post_mask, B, C, F, T,
                                    static_cast<float2*>(enh), stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_apply");
}

static int T_used_for(const Geometry& g, int T, int n_out) {
  const int cap = (n_out + 2 * g.pad + g.hop - 1) / g.hop;
  return T < cap ? T : cap;
}

int setk_istft(setk_plan_t* pl, const void* enh, int32_t B, int32_t T, int32_t n_out, const float* norm,
               float* wave, void* stream) {
  if (!pl || !enh || !wave) return fail(SETK_EINVAL, "setk_istft: null argument");
  if (B < 1 || B > pl->cfg.max_batch || T < 1 || n_out < 1)
    return fail(SETK_ESHAPE, "setk_istft: bad shape B=%d T=%d n_out=%d", B, T, n_out);
  const Geometry& g = pl->geo;
  const int T_used = T_used_for(g, T, n_out);
  cudaError_t e = ensure(&pl->d_frames_ws, &pl->frames_ws_bytes, sizeof(float) * (size_t)B * T_used * g.n_fft);
  if (e == cudaSuccess) e = ensure(&pl->d_peak, &pl->peak_bytes, sizeof(unsigned) * 2 * (size_t)pl->cfg.max_batch);
  if (e != cudaSuccess) return cuda_fail(e, "setk_istft(workspace)");
  unsigned* peak = norm ? pl->d_peak : nullptr;
  if (peak) {
    e = cudaMemsetAsync(peak, 0, sizeof(unsigned) * B, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "setk_istft(memset)");
  }
  e = run_istft_generic(pl, static_cast<const float2*>(enh), B, T, T_used, n_out, nullptr, pl->d_frames_ws,
                        wave, peak, stream);
  if (e == cudaSuccess && norm) e = run_peak_scale(wave, B, n_out, norm, peak, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_istft");
}

// wave != null: float output (+ norm rescale in place); pcm != null: the float wave goes to a plan
// workspace and the rescale pass writes PCM-16 instead
static int apply_istft_impl(setk_plan_t* pl, const float* audio, const int32_t* n_samples, int32_t B,
                            int32_t N, const void* w, int32_t w_dtype, const float* post_mask,
                            int32_t n_out, const float* norm, float* wave, int16_t* pcm, void* stream) {
  int rc = check_batch(pl, B, N, "setk_apply_istft");
  if (rc) return rc;
  if (!audio || !w || (!wave && !pcm)) return fail(SETK_EINVAL, "setk_apply_istft: null buffer");
  if (n_out < 1) return fail(SETK_ESHAPE, "setk_apply_istft: n_out=%d", n_out);
  const Geometry& g = pl->geo;
  const int T = frames_of(N, g.n_fft, g.hop, g.pad);
  cudaError_t e = ensure(&pl->d_peak, &pl->peak_bytes, sizeof(unsigned) * 2 * (size_t)pl->cfg.max_batch);
  if (e == cudaSuccess && pcm) {
    e = ensure(&pl->d_wave_ws, &pl->wave_ws_bytes, sizeof(float) * (size_t)B * n_out);
    wave = pl->d_wave_ws;
  }
  if (e != cudaSuccess) return cuda_fail(e, "setk_apply_istft(workspace)");
  unsigned* peak = norm ? pl->d_peak : nullptr;
  if (peak) {
    e = cudaMemsetAsync(peak, 0, sizeof(unsigned) * B, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "setk_apply_istft(memset)");
  }
  if (apply_istft_fused_supported(g)) {
    if (n_samples)
      e = ensure(&pl->d_tile_prefix, &pl->tile_prefix_bytes, sizeof(int) * ((size_t)pl->cfg.max_batch + 1));
    if (e != cudaSuccess) return cuda_fail(e, "setk_apply_istft(workspace)");
    e = run_apply_istft_fused(pl, audio, n_samples, B, N, T, w, w_dtype, post_mask, n_out,
                              pl->d_tile_prefix, wave, peak, stream);
  } else if (stft_spill_supported(g)) {
    // tile STFT into the bin-major workspace, y = w^H x over it, then the
    // frame-wise inverse FFT + overlap-add reading Y[b][t][f] in place
    const int T_used = T_used_for(g, T, n_out);
    const int groups = (g.C + 3) / 4;
    const int chunks = stft_cov_pick_chunks(pl, B * groups, T);
    const long long pitch = (long long)(apply_spill_bytes(g, 1, 1) / sizeof(float2));
    e = ensure(&pl->d_stft_ws, &pl->stft_ws_bytes, stft_spill_bytes(g, B, T));
    if (e == cudaSuccess) e = ensure(&pl->d_enh_ws, &pl->enh_ws_bytes, apply_spill_bytes(g, B, T));
    if (e == cudaSuccess)
      e = ensure(&pl->d_frames_ws, &pl->frames_ws_bytes, sizeof(float) * (size_t)B * T_used * g.n_fft);
    if (e == cudaSuccess)
      e = run_stft_spill(pl, audio, n_samples, B, N, T, chunks, pl->d_stft_ws, nullptr, stream);
    if (e == cudaSuccess)
      e = run_apply_spill(pl, pl->d_stft_ws, w, w_dtype, post_mask, B, T, pl->d_enh_ws, stream);
    if (e == cudaSuccess)
      e = run_istft_strided(pl, pl->d_enh_ws, (long long)T * pitch, 1, pitch, B, T_used, n_out, n_samples,
                            pl->d_frames_ws, wave, peak, stream);
  } else {
    const int T_used = T_used_for(g, T, n_out);
    e = ensure(&pl->d_stft_ws, &pl->stft_ws_bytes, sizeof(float2) * (size_t)B * g.C * g.F * T);
    if (e == cudaSuccess) e = ensure(&pl->d_enh_ws, &pl->enh_ws_bytes, sizeof(float2) * (size_t)B * g.F * T);
    if (e == cudaSuccess)
      e = ensure(&pl->d_frames_ws, &pl->frames_ws_bytes, sizeof(float) * (size_t)B * T_used * g.n_fft);
    if (e == cudaSuccess) e = run_stft_generic(pl, audio, n_samples, B, N, T, pl->d_stft_ws, stream);
    if (e == cudaSuccess)
      e = run_apply_generic(pl->d_stft_ws, w, w_dtype, post_mask, B, g.C, g.F, T, pl->d_enh_ws, stream);
    if (e == cudaSuccess)
      e = run_istft_generic(pl, pl->d_enh_ws, B, T, T_used, n_out, n_samples, pl->d_frames_ws, wave, peak,
                            stream);
  }
  if (e == cudaSuccess && pcm) e = run_peak_scale_pcm16(wave, B, n_out, norm, peak, pcm, stream);
  else if (e == cudaSuccess && norm) e = run_peak_scale(wave, B, n_out, norm, peak, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_apply_istft");
}

int setk_apply_istft(setk_plan_t* pl, const float* audio, const int32_t* n_samples, int32_t B, int32_t N,
                     const void* w, int32_t w_dtype, const float* post_mask, int32_t n_out,
                     const float* norm, float* wave, void* stream) {
  if (!wave) return fail(SETK_EINVAL, "setk_apply_istft: null buffer");
  return apply_istft_impl(pl, audio, n_samples, B, N, w, w_dtype, post_mask, n_out, norm, wave, nullptr,
                          stream);
}

int setk_apply_istft_pcm16(setk_plan_t* pl, const float* audio, const int32_t* n_samples, int32_t B,
                           int32_t N, const void* w, int32_t w_dtype, const float* post_mask,
                           int32_t n_out, const float* norm, int16_t* pcm, void* stream) {
  if (!pcm) return fail(SETK_EINVAL, "setk_apply_istft_pcm16: null buffer");
  return apply_istft_impl(pl, audio, n_samples, B, N, w, w_dtype, post_mask, n_out, norm, nullptr, pcm,
                          stream);
}

int setk_cgmm_masks(setk_plan_t* pl, const float* audio, const int32_t* n_samples, int32_t B, int32_t N,
                    int32_t num_classes, int32_t num_iters, const float* init_gamma, int32_t update_alpha,
                    float* masks, uint32_t* status, void* stream) {
  int rc = check_batch(pl, B, N, "setk_cgmm_masks");
  if (rc) return rc;
  if (!audio || !masks) return fail(SETK_EINVAL, "setk_cgmm_masks: null buffer");
  if (num_classes < 2 || num_classes > 4)
    return fail(SETK_EINVAL, "setk_cgmm_masks: num_classes %d outside 2..4", num_classes);
  if (!init_gamma && num_classes != 2)
    return fail(SETK_EINVAL, "setk_cgmm_masks: %d classes need init_gamma (the reference's start is random)",
                num_classes);
  if (num_iters < 0) return fail(SETK_EINVAL, "setk_cgmm_masks: num_iters %d", num_iters);
  const Geometry& g = pl->geo;
  if (!stft_spill_supported(g))
    return fail(SETK_EUNSUPPORTED, "setk_cgmm_masks: needs n_fft 512 or 1024 (got %d, hop %d)", g.n_fft, g.hop);
  const int T = frames_of(N, g.n_fft, g.hop, g.pad);
  if (T < 1) return fail(SETK_ESHAPE, "setk_cgmm_masks: no frame in %d samples", N);
  const int P = (int)(apply_spill_bytes(g, 1, 1) / sizeof(float2));
  CgCtx ctx;
  ctx.geo = g; ctx.sm_count = pl->sm_count;
  cudaError_t e = ensure(&pl->d_stft_ws, &pl->stft_ws_bytes, stft_spill_bytes(g, B, T));
  if (e == cudaSuccess)
    e = ensure(&pl->d_cgmm_ws, &pl->cgmm_ws_bytes, cgmm_workspace_bytes(&ctx, B, T, num_classes, P));
  if (e != cudaSuccess) return cuda_fail(e, "setk_cgmm_masks(workspace)");
  const int groups = (g.C + 3) / 4;
  e = run_stft_spill(pl, audio, n_samples, B, N, T, stft_cov_pick_chunks(pl, B * groups, T), pl->d_stft_ws,
                     nullptr, stream);
  if (e == cudaSuccess)
    e = run_cgmm(&ctx, pl->d_stft_ws, P, pl->d_cgmm_ws, n_samples, B, N, T, num_classes, num_iters, init_gamma,
                 update_alpha, masks, status, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_cgmm_masks");
}

int setk_cgmm_stft(const void* stft, int32_t B, int32_t C, int32_t F, int32_t T, int32_t num_classes,
                   int32_t num_iters, const float* init_gamma, int32_t update_alpha, float* masks,
                   uint32_t* status, void* stream) {
  if (!stft || !masks) return fail(SETK_EINVAL, "setk_cgmm_stft: null buffer");
  if (B < 1 || B > 65535 || F < 1 || T < 1 || C < 1 || C > SETK_MAX_CHANNELS)
    return fail(SETK_ESHAPE, "setk_cgmm_stft: bad shape B=%d C=%d F=%d T=%d", B, C, F, T);
  if (num_classes < 2 || num_classes > 4)
    return fail(SETK_EINVAL, "setk_cgmm_stft: num_classes %d outside 2..4", num_classes);
  if (!init_gamma && num_classes != 2)
    return fail(SETK_EINVAL, "setk_cgmm_stft: %d classes need init_gamma (the reference's start is random)",
                num_classes);
  if (num_iters < 0) return fail(SETK_EINVAL, "setk_cgmm_stft: num_iters %d", num_iters);
  CgCtx ctx;
  // frames are given, not derived from samples: a geometry whose frame count never limits
  ctx.geo.C = C; ctx.geo.F = F; ctx.geo.n_fft = 2; ctx.geo.log2n = 1; ctx.geo.hop = 1; ctx.geo.pad = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx.sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return cuda_fail(e, "setk_cgmm_stft(device)");
  const int P = (F + 7) & ~7;
  const size_t x_bytes = sizeof(float2) * (size_t)B * T * C * P;
  const size_t w_bytes = cgmm_workspace_bytes(&ctx, B, T, num_classes, P);
  char* ws = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  e = cudaMallocAsync(reinterpret_cast<void**>(&ws), x_bytes + w_bytes + 256, st);   // stream-ordered scratch
  if (e != cudaSuccess) return cuda_fail(e, "setk_cgmm_stft(workspace)");
  float2* X = reinterpret_cast<float2*>(ws);
  double* cw = reinterpret_cast<double*>(ws + ((x_bytes + 255) / 256) * 256);
  e = run_bcft_to_spill(static_cast<const float2*>(stft), B, C, F, T, P, X, stream);
  if (e == cudaSuccess)
    e = run_cgmm(&ctx, X, P, cw, nullptr, B, /*N=*/T + 1, T, num_classes, num_iters, init_gamma, update_alpha,
                 masks, status, stream);
  cudaError_t ef = cudaFreeAsync(ws, st);
  if (e == cudaSuccess) e = ef;
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_cgmm_stft");
}

static int wpe_entry(const char* who, const void* stft, const void* lambda_enh, int32_t B, int32_t C, int32_t F,
                     int32_t T, int32_t taps, int32_t delay, int32_t context, int32_t num_iters, void* out,
                     float* inv_lambda, uint32_t* status, void* stream) {
  if (!stft || !out) return fail(SETK_EINVAL, "%s: null buffer", who);
  if (B < 1 || F < 1 || T < 1 || C < 1 || C > SETK_MAX_CHANNELS || (long long)B * F > 2000000000LL)
    return fail(SETK_ESHAPE, "%s: bad shape B=%d C=%d F=%d T=%d", who, B, C, F, T);
  if (taps < 1 || delay < 0 || context < 0 || num_iters < 1)
    return fail(SETK_EINVAL, "%s: taps=%d delay=%d context=%d num_iters=%d", who, taps, delay, context,
                num_iters);
  if (!wpe_supported(C, T, taps, delay, context))
    return fail(SETK_EUNSUPPORTED, "%s: %d channels x %d taps (context %d) exceeds the kernels' "
                "shared-memory budget", who, C, taps, context);
  const int P = (F + 7) & ~7;
  const size_t x_bytes = sizeof(float2) * (size_t)B * T * C * P;
  const size_t w_bytes = wpe_workspace_bytes(B, C, F, taps);
  char* ws = nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&ws), x_bytes + w_bytes + 256, st);
  if (e != cudaSuccess) return cuda_fail(e, who);
  float2* X = reinterpret_cast<float2*>(ws);
  double* dw = reinterpret_cast<double*>(ws + ((x_bytes + 255) / 256) * 256);
  e = run_bcft_to_spill(static_cast<const float2*>(stft), B, C, F, T, P, X, stream);
  if (e == cudaSuccess)
    e = run_wpe(X, P, B, C, F, T, taps, delay, context, num_iters, dw, static_cast<float2*>(out), status,
                static_cast<const float2*>(lambda_enh), inv_lambda, stream);
  cudaError_t ef = cudaFreeAsync(ws, st);
  if (e == cudaSuccess) e = ef;
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, who);
}

int setk_wpe_stft(const void* stft, int32_t B, int32_t C, int32_t F, int32_t T, int32_t taps, int32_t delay,
                  int32_t context, int32_t num_iters, void* out, uint32_t* status, void* stream) {
  return wpe_entry("setk_wpe_stft", stft, nullptr, B, C, F, T, taps, delay, context, num_iters, out, nullptr,
                   status, stream);
}

int setk_wpe_step(const void* stft, const void* lambda_enh, int32_t B, int32_t C, int32_t F, int32_t T,
                  int32_t taps, int32_t delay, int32_t context, void* out, float* inv_lambda,
                  uint32_t* status, void* stream) {
  return wpe_entry("setk_wpe_step", stft, lambda_enh, B, C, F, T, taps, delay, context, 1, out, inv_lambda,
                   status, stream);
}

int setk_float_to_pcm16(const float* wave, int64_t n, int16_t* pcm, void* stream) {
  if (!wave || !pcm || n < 0) return fail(SETK_EINVAL, "setk_float_to_pcm16: bad argument");
  if (n == 0) return SETK_OK;
  cudaError_t e = run_float_to_pcm16(wave, n, pcm, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_float_to_pcm16");
}

int setk_pcm16_to_float(const int16_t* pcm, int64_t n, float* wave, void* stream) {
  if (!wave || !pcm || n < 0) return fail(SETK_EINVAL, "setk_pcm16_to_float: bad argument");
  if (n == 0) return SETK_OK;
  cudaError_t e = run_pcm16_to_float(pcm, n, wave, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_pcm16_to_float");
}

int setk_cm_masks(const uint8_t* blobs, int64_t slot_bytes, int32_t B, int32_t T, int32_t F, float* out,
                  int32_t* status, void* stream) {
  if (!blobs || !out || !status || B < 0 || T < 0 || F < 0 || slot_bytes < 16 || (slot_bytes & 15) ||
      (reinterpret_cast<uintptr_t>(blobs) & 15))
    return fail(SETK_EINVAL, "setk_cm_masks: bad argument");
  cudaError_t e = run_cm_masks(blobs, slot_bytes, B, T, F, out, status, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_cm_masks");
}

// ---- spatial features (spatial.cu) ----
int setk_ipd(const void* si, const void* sj, int64_t rows, int32_t F, int32_t mode, float* out,
             void* stream) {
  if (!si || !sj || !out) return fail(SETK_EINVAL, "setk_ipd: null buffer");
  if (rows < 1 || F < 1) return fail(SETK_ESHAPE, "setk_ipd: rows=%lld F=%d", (long long)rows, F);
  if (mode < 0 || mode > 2) return fail(SETK_EINVAL, "setk_ipd: mode %d outside 0..2", mode);
  cudaError_t e = run_ipd(static_cast<const float2*>(si), static_cast<const float2*>(sj), rows, F, mode,
                          out, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_ipd");
}

int setk_directional_feats(const void* stft, const void* steer, int32_t steer_batched,
                           const int32_t* pairs, int32_t n_pairs, int32_t B, int32_t M, int32_t F,
                           int32_t T, double* out, void* stream) {
  if (!stft || !steer || !out) return fail(SETK_EINVAL, "setk_directional_feats: null buffer");
  if (B < 1 || F < 1 || T < 1 || M < 2 || M > SETK_MAX_CHANNELS)
    return fail(SETK_ESHAPE, "setk_directional_feats: bad shape B=%d M=%d F=%d T=%d", B, M, F, T);
  if (pairs && n_pairs < 1) return fail(SETK_EINVAL, "setk_directional_feats: empty pair list");
  cudaError_t e = run_dirfeat(static_cast<const float2*>(stft), static_cast<const double2*>(steer),
                              steer_batched, pairs, n_pairs, B, M, F, T, out, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_directional_feats");
}

int64_t setk_gcc_phat_work_doubles(int32_t T, int32_t F, int32_t D) {
  if (T < 1 || F < 1 || D < 1) return -1;
  return (int64_t)gcc_phat_work_doubles(T, F, D);
}

int setk_gcc_phat(const void* si, const void* sj, int32_t T, int32_t F, const double* omega,
                  const double* tau, int32_t D, int32_t normalize, int32_t apply_floor,
                  int32_t accumulate, double* work, double* out, void* stream) {
  if (!si || !sj || !omega || !tau || !work || !out) return fail(SETK_EINVAL, "setk_gcc_phat: null buffer");
  if (T < 1 || F < 1 || D < 1) return fail(SETK_ESHAPE, "setk_gcc_phat: T=%d F=%d D=%d", T, F, D);
  if ((size_t)F * sizeof(double2) > 48 * 1024)
    return fail(SETK_EUNSUPPORTED, "setk_gcc_phat: F=%d too large (<= 3072 bins)", F);
  cudaError_t e = run_gcc_phat(static_cast<const float2*>(si), static_cast<const float2*>(sj), T, F, omega,
                               tau, D, normalize, apply_floor, accumulate, work, out, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_gcc_phat");
}

int64_t setk_msc_work_doubles(int32_t T, int32_t F) {
  if (T < 1 || F < 1) return -1;
  return (int64_t)msc_work_doubles(T, F);
}

int setk_msc(const void* spec, int32_t N, int32_t T, int32_t F, int32_t context, int32_t normalize,
             double* work, double* out, void* stream) {
  if (!spec || !work || !out) return fail(SETK_EINVAL, "setk_msc: null buffer");
  if (N < 2 || N > SETK_MAX_CHANNELS || T < 1 || F < 1)
    return fail(SETK_ESHAPE, "setk_msc: bad shape N=%d T=%d F=%d", N, T, F);
  if (context < 0) return fail(SETK_EINVAL, "setk_msc: context %d < 0", context);
  cudaError_t e = run_msc(static_cast<const float2*>(spec), N, T, F, context, normalize, work, out, stream);
  return e == cudaSuccess ? SETK_OK : cuda_fail(e, "setk_msc");
}

}  // extern "C"
