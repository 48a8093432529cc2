// apply_istft_ws.cu -- beamform apply fused with the inverse STFT, warp-specialised build
// (round 2) of apply_istft_fused.cu for the metric geometry: 4 channels, 512-point frames,
// hop 256.  Same maths, schedule (TileSched) and outputs; other geometries keep the classic kernel.
//
// Replaces (scripts/sptk): Beamformer.beamform (libs/beamformer.py:220-234), the optional
// post-mask (apply_adaptive_beamformer.py:174-175) and inverse_stft (libs/utils.py:142-173 ->
// librosa.istft: irfft, x window, overlap-add, / window-sum-square, trim) up to the peak of |y|
// needed by the `norm` rescale.  C++ twin: Beamform (include/beamformer.cc:215-230) +
// InverseShortTimeFT (include/stft.cc:154-198).
//
// 512 threads = two FFT groups of 4 warps (64 registers) + 4 IFFT warps (48) + 4 BACK warps (80,
// the highest warp ids), launched at 64 => two CTAs per SM (32 warps).  History (B200, config 2,
// classic kernel 0.57 ms): a first build let two of the FFT warps run the inverse transforms --
// they waited for the BACK warps' apply while the other six waited for them at the audio barrier,
// the roles ran one after the other: 0.79 ms (profiles/r2_apply_istft_ws_serialised_ncu.txt);
// inverse-FFT warps of their own: 0.62 ms, still bound by the ring: a tile passes four stages
// (forward FFT -> apply -> inverse FFT -> flush) but only TWO slots existed, so a stage could start
// only every (sum of the four) / 2.  This build cuts tiles to 2 frames x 4 channels and rings FOUR
// slots in the same shared memory, one per stage; the two FFT groups take alternate tiles with an
// audio buffer each, so nothing but mbarriers couples the stages.  MEASURED: 0.70 ms -- slower
// again (profiles/r2_apply_istft_ws_ring4_ncu.txt): twice the tiles means twice the per-tile
// protocol (waits, arrivals, table reads; 53 % of the issued instructions are mbarrier polls) and
// a table refill every 128 tiles drains the four-stage pipeline.  The classic kernel therefore
// stays the default (apply_istft_fused.cu, SETK_AI_IMPL=ws selects this one); what the three
// builds show is that this pass has too many short stages per frame for a producer/consumer ring
// to pay -- the next attempt should fuse apply + inverse FFT in one role.  A slot is re-used in place:
//
//   FFT group   audio tile (TMA bulk copy) -> forward FFT -> Z[slot]        ... arrive z_full
//   BACK warps  wait z_full: thread k applies the weights of bin pair (k, 256-k) to the four
//               channels of each frame and writes the half-size inverse spectrum Zi over
//               channel 0's part of the slot (it owns entries k, 256-k)     ... arrive zi_full
//   IFFT warp   (tile mod 4) wait zi_full: inverse FFT of the two frames, x synthesis window
//               -> frames over channel 1's part                             ... arrive fr_full
//   BACK warps  wait fr_full (of the PREVIOUS tile, so the inverse FFT overlaps their apply):
//               overlap-add with the carried half frame, / window-sum-square, trim, 16-byte
//               stores, running peak                                        ... arrive z_empty
//
// The CTA's tiles -- including the one-frame halo tile a run needs in front of its first frame
// for the overlap-add -- are described once per tile in a small table (AwTile) filled by one
// thread per 128 tiles, instead of every thread re-deriving the schedule on every tile.
// Deterministic: no atomics on data (the peak is an order-independent max).
// Algorithmic bytes per utterance: 4*C*N (audio) + 8*F*C (w) + 4*n_out (wave).
#include <cstdlib>
#include "common.cuh"
#include "stft_tile.cuh"
#include "apply_istft_args.cuh"

namespace setk {

#define SETK_AW_LAUNCH_REGS 64
#define SETK_AW_IFFT_REGS 48
#define SETK_AW_BACK_REGS 80

constexpr int kAwFftThreads = 256, kAwIfftThreads = 128, kAwBackThreads = 128, kAwThreads = 512;
constexpr int kAwGroupThreads = 128;           // one FFT group (4 warps = the 8 half-warp jobs of a tile)
constexpr int kAwBarFft = 1, kAwBarBack = 3;   // named barriers 1, 2: the FFT groups; 3: BACK
constexpr int kAwSlots = 4, kAwLook = 2;       // ring slots; look-ahead entries of the tile table
#ifndef SETK_TABLE_CHUNK
#define SETK_TABLE_CHUNK 128       // tile descriptors per table fill (the CPU test tier builds with 8)
#endif
constexpr int kAwChunk = SETK_TABLE_CHUNK;
constexpr int kAwC = 4, kAwTT = 2;
constexpr int kAwJobs = kAwC * kAwTT;         // half-warp FFT jobs (and Z parts) per slot
constexpr int kAwWPitch = 260;

enum : unsigned {
  AW_NT = 0xfu,
  AW_AUDIO_BULK = 1u << 4,
  AW_SEG_BEGIN = 1u << 5,     // first tile (halo included) of an utterance's part of the run
  AW_SEG_END = 1u << 6,       // last tile of that part: peak flushed
  AW_UTT_END = 1u << 7,       // ... and the utterance's own last frames are in it: tail zero-filled
};
struct alignas(16) AwTile { int b, t0, t_begin; unsigned flags; };

struct AwGen {            // generator of the CTA's tile sequence (one thread)
  int x;                  // next linear tile of the launch's schedule
  int halo_done;          // the halo tile of the segment starting at x has been emitted
  int done;               // nothing left
  int count;              // entries of the current table fill (consumers read this)
  int avail;              // count + valid look-ahead entries behind them
  int pad_[3];
};

struct AwSmem {
  AwTile* tiles;          // [kAwChunk + kAwLook]
  AwGen* gen;
  MBar* bar_audio;        // [2] one per FFT group
  MBar* z_full;           // [4] 4 warps of an FFT group
  MBar* zi_full;          // [4] 4 BACK warps
  MBar* fr_full;          // [4] 1 IFFT warp
  MBar* z_empty;          // [4] 4 BACK warps
  float* win;             // [512] analysis window x 0.5
  float2* twtab;          // [256]
  float* wsyn;            // [512] window / 512
  float* rw;              // [256] 1 / (wsq[r] + wsq[r + 256])
  float* carry;           // [2][256]
  float2* w;              // [4][kAwWPitch]
  float* audio;           // [2 groups][4][Lp]
  float2* z;              // [kAwSlots][kAwJobs][SETK_ZSLOT]
  int Lp;
  SETK_HD static int staged_len(int hop) { return ((kAwTT - 1) * hop + kNfft + 3) & ~3; }
  SETK_HD static size_t bytes(int hop) {
    return 256 + sizeof(AwTile) * (kAwChunk + kAwLook + 2) + sizeof(float) * (2 * kNfft + kM + 2 * kM) +
           sizeof(float2) * (256 + kAwC * kAwWPitch) + sizeof(float) * 2 * kAwC * staged_len(hop) +
           sizeof(float2) * kAwSlots * kAwJobs * SETK_ZSLOT;
  }
  __device__ void carve(float* base, int hop) {
    Lp = staged_len(hop);
    MBar* bars = reinterpret_cast<MBar*>(base);
    bar_audio = bars; z_full = bars + 2; zi_full = bars + 6; fr_full = bars + 10; z_empty = bars + 14;
    gen = reinterpret_cast<AwGen*>(bars + 20);                 // byte 160 .. 192
    tiles = reinterpret_cast<AwTile*>(base + 64);              // byte 256
    win = base + 64 + 4 * (kAwChunk + kAwLook + 2);
    wsyn = win + kNfft;
    rw = wsyn + kNfft;
    carry = rw + kM;
    twtab = reinterpret_cast<float2*>(carry + 2 * kM);
    w = twtab + 256;
    audio = reinterpret_cast<float*>(w + kAwC * kAwWPitch);
    z = reinterpret_cast<float2*>(audio + 2 * kAwC * Lp);
  }
};

// frames librosa.istft uses of utterance b for this output length
__device__ __forceinline__ int aw_frames_used(const ApplyIstftArgs& a, int b, int T_cap, int& nb) {
  nb = a.n_samples ? a.n_samples[b] : a.N;
  return imin(frames_of(nb, kNfft, a.g.hop, a.g.pad), T_cap);
}

// ONE thread: the next `want` tiles of this CTA's run [lo, hi) -> table (and one look-ahead entry)
__device__ void aw_fill_table(const ApplyIstftArgs& a, const AwSmem& sm, int lo, int hi, int T_cap,
                              bool vec_ok) {
  AwGen g = *sm.gen;
  int n = 0;
  AwGen before_last = g;
  while (n < kAwChunk + kAwLook && !g.done) {
    if (n == kAwChunk) before_last = g;          // the look-ahead entries are generated again next time
    const int x = g.x;
    const int b = sched_find(a.sched, x);
    const int pb = sched_prefix(a.sched, b), pe = sched_prefix(a.sched, b + 1);
    const int seg_lo = imax(lo, pb), seg_hi = imin(hi, pe);
    int nb;
    const int T_used = aw_frames_used(a, b, T_cap, nb);
    const int t_begin = (seg_lo - pb) * kAwTT;
    const int t_end = imin((seg_hi - pb) * kAwTT, T_used);
    AwTile d;
    d.b = b; d.t_begin = t_begin;
    unsigned f = 0;
    const bool first = x == seg_lo;
    if (first && !g.halo_done && t_begin > 0 && t_begin < t_end) {
      // the frame before the run's first one: its second half overlaps the first owned row
      d.t0 = t_begin - 1;
      f = 1u | AW_SEG_BEGIN;
      if (tile_bulk_ok(d.t0, 1, a.g.hop, a.g.pad, nb, vec_ok)) f |= AW_AUDIO_BULK;
      g.halo_done = 1;
    } else {
      d.t0 = (x - pb) * kAwTT;
      const int nt = imax(0, imin(kAwTT, t_end - d.t0));
      f = (unsigned)nt;
      if (nt > 0 && tile_bulk_ok(d.t0, nt, a.g.hop, a.g.pad, nb, vec_ok)) f |= AW_AUDIO_BULK;
      if (first && !g.halo_done) f |= AW_SEG_BEGIN;
      if (x + 1 == seg_hi) {
        f |= AW_SEG_END;
        if ((seg_hi - pb) * kAwTT >= T_used) f |= AW_UTT_END;
      }
      g.halo_done = 0;
      g.x = x + 1;
      if (g.x >= hi) g.done = 1;
    }
    d.flags = f;
    sm.tiles[n++] = d;
  }
  int count = n;
  if (n > kAwChunk) { count = kAwChunk; g = before_last; }   // entries >= kAwChunk are look-ahead only
  g.count = count;
  g.avail = n;
  *sm.gen = g;
}

__device__ __forceinline__ void aw_stage_bulk(float* dst, int Lp, MBar* bar, const float* __restrict__ xb,
                                              int N, int t0, int nt, int hop, int pad) {   // ONE thread
  const int need = (nt - 1) * hop + kNfft;
  const int i0 = t0 * hop - pad;
  fence_proxy_async();
  mbar_expect_tx(bar, (unsigned)(kAwC * need * sizeof(float)));
#pragma unroll
  for (int c = 0; c < kAwC; ++c)
    bulk_g2s(dst + c * Lp, xb + (long long)c * N + i0, (unsigned)(need * sizeof(float)), bar);
}
__device__ __noinline__ void aw_stage_scalar(float* dst, int Lp, const float* __restrict__ xb, int N, int nb,
                                             int t0, int nt, int hop, int pad, int gtid) {
  const int p0 = t0 * hop;
  const int need = (nt - 1) * hop + kNfft;
  for (int c = 0; c < kAwC; ++c) {
    const float* src = xb + (long long)c * N;
    for (int q = gtid; q < need; q += kAwGroupThreads) {
      const int i = pad ? reflect_index(p0 + q, pad, nb) : (p0 + q);
      dst[c * Lp + q] = src[i];
    }
  }
}

// ---------------------------------------------------------------------------
// FFT role: threads 0..255 = two groups of 4 warps; group g transforms the tiles n = g (mod 2);
// half-warp job = (thread in group) / 16 = frame * 4 + channel
// ---------------------------------------------------------------------------
__device__ __forceinline__ void aw_fft_role(const ApplyIstftArgs& a, const AwSmem& sm, int lo, int hi,
                                            int T_cap, bool vec_ok) {
  const int grp = (int)threadIdx.x >> 7, gtid = (int)threadIdx.x & (kAwGroupThreads - 1);
  const int lane = gtid & 31, lane16 = lane & 15;
  const int job = gtid >> 4;
  const int fr = job >> 2, ch = job & 3;
  const int hop = a.g.hop, pad = a.g.pad;
  float* abuf = sm.audio + grp * kAwC * sm.Lp;
  MBar* abar = sm.bar_audio + grp;
  unsigned apar = 0;

  auto stage = [&](const AwTile& d) -> bool {        // true: written with ordinary stores
    const int nt = (int)(d.flags & AW_NT);
    if (nt <= 0) return false;
    const float* xb = a.audio + ((long long)d.b * a.c_total + a.c0) * a.N;
    if (d.flags & AW_AUDIO_BULK) {
      if (gtid == 0) aw_stage_bulk(abuf, sm.Lp, abar, xb, a.N, d.t0, nt, hop, pad);
      return false;
    }
    const int nb = a.n_samples ? a.n_samples[d.b] : a.N;
    aw_stage_scalar(abuf, sm.Lp, xb, a.N, nb, d.t0, nt, hop, pad, gtid);
    return true;
  };

  int nbase = 0;                                     // local number of the table's first tile
  bool first_fill = true;
  for (;;) {
    __syncthreads();                                 // the previous table is consumed by every role
    if (threadIdx.x == 0) aw_fill_table(a, sm, lo, hi, T_cap, vec_ok);
    __syncthreads();
    const int cnt = sm.gen->count, avail = sm.gen->avail;
    const bool more = !sm.gen->done;
    if (first_fill) {
      first_fill = false;
      if (grp < avail && stage(sm.tiles[grp])) named_bar_sync(kAwBarFft + grp, kAwGroupThreads);
    }
    for (int i = grp; i < cnt; i += 2) {             // kAwChunk is even: i and n have the same parity
      const int n = nbase + i;
      const AwTile d = sm.tiles[i];
      const int nt = (int)(d.flags & AW_NT);
      float2 v[16];
      if (nt > 0) {
        if (d.flags & AW_AUDIO_BULK) { mbar_wait(abar, apar); apar ^= 1u; }
        const int fr_src = imin(fr, nt - 1);         // a dead frame re-transforms the last live one
        const float* src = abuf + ch * sm.Lp + fr_src * hop + 2 * lane16;
        const float* wsrc = sm.win + 2 * lane16;
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) {
          const float2 sx = *reinterpret_cast<const float2*>(src + 32 * m1);
          const float2 w = *reinterpret_cast<const float2*>(wsrc + 32 * m1);
          v[m1] = f2mul(sx, w);
        }
      }
      named_bar_sync(kAwBarFft + grp, kAwGroupThreads);   // the group holds its samples: its buffer is free
      bool scalar_next = false;
      if (i + 2 < avail) scalar_next = stage(sm.tiles[i + 2]);   // this group's next tile
      if (nt > 0) halfwarp_fft256_a(v, sm.twtab, lane16);
      const int s = n & (kAwSlots - 1);
      mbar_wait(&sm.z_empty[s], ((unsigned)(n >> 2) & 1u) ^ 1u);   // tile n - 4 has been flushed
      if (nt > 0) {
        float2* zs = sm.z + (s * kAwJobs + job) * SETK_ZSLOT;
        halfwarp_fft256_b(v, zs, lane16);
#pragma unroll
        for (int p = 0; p < 16; ++p) zs[lane16 + 16 * kof(p)] = v[p];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.z_full[s]);
      if (scalar_next) named_bar_sync(kAwBarFft + grp, kAwGroupThreads);
    }
    nbase += cnt;
    if (!more) break;
  }
}

// ---------------------------------------------------------------------------
// IFFT role: threads 256..383; warp w takes the tiles n = w (mod 4), a half-warp per frame:
// conj . FFT256 . conj of the half-size inverse spectrum, x synthesis window
// ---------------------------------------------------------------------------
__device__ __forceinline__ void aw_ifft_role(const ApplyIstftArgs& a, const AwSmem& sm, int lo, int hi,
                                             int T_cap, bool vec_ok) {
  const int tid = (int)threadIdx.x - kAwFftThreads;
  const int lane = tid & 31, lane16 = lane & 15, iw = tid >> 5;
  int nbase = 0;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) aw_fill_table(a, sm, lo, hi, T_cap, vec_ok);
    __syncthreads();
    const int cnt = sm.gen->count;
    const bool more = !sm.gen->done;
    for (int i = 0; i < cnt; ++i) {
      const int n = nbase + i;
      if ((n & (kAwSlots - 1)) != iw) continue;      // another warp's tile
      const int nt_m = (int)(sm.tiles[i].flags & AW_NT);
      const int s = n & (kAwSlots - 1);
      mbar_wait(&sm.zi_full[s], (unsigned)(n >> 2) & 1u);
      const int j = lane >> 4;                       // frame inside the tile
      {
        // a dead frame (j >= nt_m) transforms zeros: both half-warps run the same code
        float2* zi = sm.z + (s * kAwJobs + j * kAwC) * SETK_ZSLOT;
        float2 v[16];
        const bool live = j < nt_m;
#pragma unroll
        for (int m1 = 0; m1 < 16; ++m1) {
          const float2 x = live ? zi[16 * m1 + lane16] : make_float2(0.f, 0.f);
          v[m1] = make_float2(x.x, -x.y);
        }
        __syncwarp();                                // the slot part is free: reuse it as the exchange tile
        halfwarp_fft256_a(v, sm.twtab, lane16);
        halfwarp_fft256_b(v, zi, lane16);
        float* fo = reinterpret_cast<float*>(sm.z + (s * kAwJobs + j * kAwC + 1) * SETK_ZSLOT);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const int mm = lane16 + 16 * kof(q);
          const float2 ws = *reinterpret_cast<const float2*>(sm.wsyn + 2 * mm);
          *reinterpret_cast<float2*>(fo + 2 * mm) = make_float2(v[q].x * ws.x, -v[q].y * ws.y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.fr_full[s]);
    }
    nbase += cnt;
    if (!more) break;
  }
}

// ---------------------------------------------------------------------------
// BACK role: threads 384..511 -- apply (bin pair per thread) and flush
// ---------------------------------------------------------------------------
struct AwSeg {            // what the flush of a tile needs to know about its utterance
  int b, T_used, own_begin, n_lim, expected;
};

__device__ __forceinline__ void aw_back_role(const ApplyIstftArgs& a, const AwSmem& sm, int lo, int hi,
                                             int T_cap, bool vec_ok) {
  constexpr int C = kAwC, TT = kAwTT, F = kBins;
  const int tid = (int)threadIdx.x - kAwFftThreads - kAwIfftThreads, lane = tid & 31, warp = tid >> 5;
  const int hop = a.g.hop, pad = a.g.pad;
  const int k = tid, km = kM - tid;
  const float2 tw = split_twiddle(tid);
  const bool out_vec = ((a.n_out & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.wave) & 15) == 0);
  float2 wk[C], wm[C];
  float peak = 0.f;
  int cur = 0;                                       // which carry buffer is the input

  // one (pair, frame) item: split Z -> X_c[k], X_c[256-k]; y = w^H x; inverse split -> Zi
  auto item = [&](float2* zt, int j, int kk, int kkm, float2 tww, const float2* wkk, const float2* wmm,
                  int b, int t0, bool post) {
    float2 yk = make_float2(0.f, 0.f), ym = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float2* z = zt + (j * C + c) * SETK_ZSLOT;
      float2 xk, xm;
      split_pair(z[kk & (kM - 1)], z[kkm & (kM - 1)], tww, xk, xm);
      if (kk == 0) { xk.y = 0.f; xm.y = 0.f; }
      yk = cmad_conjw(wkk[c], xk, yk);               // += conj(w) x: two packed instructions
      ym = cmad_conjw(wmm[c], xm, ym);
    }
    if (post) {
      const float* pm = a.post_mask + ((long long)b * a.T + (t0 + j)) * F;
      const float mk = pm[kk], mm = pm[kkm];
      yk = f2mul(yk, make_float2(mk, mk)); ym = f2mul(ym, make_float2(mm, mm));
    }
    float2* zi = zt + (j * C) * SETK_ZSLOT;          // over channel 0's part (entries kk, kkm are ours)
    if (kk == 0) {
      zi[0] = make_float2(yk.x + ym.x, yk.x - ym.x); // irfft ignores Im Y[0], Im Y[256]
    } else {
      const float2 cm = make_float2(ym.x, -ym.y);
      const float2 e = f2add(yk, cm), d = f2sub(yk, cm);
      const float2 p = cmul_conj(d, tww);            // conj(tw) D
      zi[kk] = f2add(e, p);
      if (kk != kM / 2) {
        const float2 qq = f2sub(e, p);
        zi[kkm] = make_float2(qq.x, -qq.y);
      }
    }
  };

  // overlap-add, normalise, trim, write: the frames of tile (p_t0, p_nt) in slot s
  auto flush = [&](int s, int p_t0, int p_nt, const AwSeg& sg, bool carry_zero) {
    const float* frames = reinterpret_cast<const float*>(sm.z + (s * kAwJobs + 1) * SETK_ZSLOT);
    constexpr int FSTRIDE = C * SETK_ZSLOT * 2;      // floats between the frames of a slot
    const int p_tile = p_t0 * hop;
    const bool last_tile = (p_t0 + p_nt == sg.T_used);
    float* yb = a.wave + (long long)sg.b * a.n_out;
    const float* cin = sm.carry + cur * kM;
    float* cout = sm.carry + (cur ^ 1) * kM;
    const bool interior = out_vec && p_nt == TT && !last_tile && p_t0 >= 1 &&
                          p_t0 + p_nt < sg.T_used && p_tile >= sg.own_begin && p_tile >= pad &&
                          p_tile - pad + p_nt * kM <= sg.n_lim;
    for (int it = tid; it < (p_nt + 1) * (kM / 4); it += kAwBackThreads) {
      const int jj = it >> 6, r = (it & 63) * 4;
      float4 val;
      if (jj >= 1) val = *reinterpret_cast<const float4*>(frames + (jj - 1) * FSTRIDE + kM + r);
      else if (carry_zero) val = make_float4(0.f, 0.f, 0.f, 0.f);
      else val = *reinterpret_cast<const float4*>(cin + r);
      if (jj < p_nt) {
        const float4 f = *reinterpret_cast<const float4*>(frames + jj * FSTRIDE + r);
        val.x += f.x; val.y += f.y; val.z += f.z; val.w += f.w;
      }
      if (interior) {
        if (jj < TT) {
          const float4 rw = *reinterpret_cast<const float4*>(sm.rw + r);
          val.x *= rw.x; val.y *= rw.y; val.z *= rw.z; val.w *= rw.w;
          *reinterpret_cast<float4*>(yb + (p_tile - pad + jj * kM + r)) = val;
          peak = fmaxf(peak, fmaxf(fmaxf(fabsf(val.x), fabsf(val.y)), fmaxf(fabsf(val.z), fabsf(val.w))));
        } else {
          *reinterpret_cast<float4*>(cout + r) = val;
        }
      } else if (jj < p_nt || last_tile) {
        const int p = p_tile + jj * kM + r;
        const int q = p - pad;
        if (p >= sg.own_begin && q >= 0 && q < sg.n_lim) {
          const int t = p_t0 + jj;                   // frame starting at this row
          float vv[4] = {val.x, val.y, val.z, val.w};
          if (t >= 1 && t < sg.T_used) {
            const float4 rw = *reinterpret_cast<const float4*>(sm.rw + r);
            vv[0] *= rw.x; vv[1] *= rw.y; vv[2] *= rw.z; vv[3] *= rw.w;
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float wss = (t < sg.T_used ? a.wsq[r + i] : 0.f) + (t >= 1 ? a.wsq[kM + r + i] : 0.f);
              if (wss > SETK_TINY32) vv[i] /= wss;
            }
          }
          if (out_vec && q + 4 <= sg.n_lim) {
            *reinterpret_cast<float4*>(yb + q) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            peak = fmaxf(peak, fmaxf(fmaxf(fabsf(vv[0]), fabsf(vv[1])), fmaxf(fabsf(vv[2]), fabsf(vv[3]))));
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (q + i < sg.n_lim) { yb[q + i] = vv[i]; peak = fmaxf(peak, fabsf(vv[i])); }
          }
        }
      } else {
        *reinterpret_cast<float4*>(cout + r) = val;
      }
    }
    cur ^= 1;
  };
  // end of an utterance's part of the run
  auto finish_segment = [&](const AwSeg& sg, unsigned flags) {
    if (flags & AW_UTT_END) {      // zero-fill what no frame reaches (fix_length padding / too short)
      const int q0 = a.n_samples ? sg.n_lim : imax(sg.expected - pad, 0);
      float* yb = a.wave + (long long)sg.b * a.n_out;
      for (int qq = q0 + tid; qq < a.n_out; qq += kAwBackThreads) yb[qq] = 0.f;
    }
    if (a.peak) {
      for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
      if (lane == 0 && peak > 0.f) atomicMax(a.peak + sg.b, __float_as_uint(peak));
      peak = 0.f;
    }
  };

  int n = 0;
  AwSeg seg = {0, 0, 0, 0, 0}, seg_prev = {0, 0, 0, 0, 0};
  AwTile prev = {0, 0, 0, 0};
  bool have_prev = false;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) aw_fill_table(a, sm, lo, hi, T_cap, vec_ok);
    __syncthreads();
    const int cnt = sm.gen->count;
    const bool more = !sm.gen->done;
    for (int i = 0; i < cnt; ++i, ++n) {
      const AwTile d = sm.tiles[i];
      const int nt = (int)(d.flags & AW_NT);
      if (d.flags & AW_SEG_BEGIN) {
        // this utterance's weights (the previous tile's flush does not read them)
        int nb;
        seg.b = d.b;
        seg.T_used = aw_frames_used(a, d.b, T_cap, nb);
        seg.expected = seg.T_used > 0 ? kNfft + hop * (seg.T_used - 1) : 0;
        seg.n_lim = a.n_samples ? imin(a.n_out, imax(seg.expected - 2 * pad, 0)) : a.n_out;
        seg.own_begin = d.t_begin * hop;
        named_bar_sync(kAwBarBack, kAwBackThreads);  // everyone has read the old weights
        for (int e = tid; e < F * C; e += kAwBackThreads) {
          const int kk = e / C, c = e - kk * C;
          const long long wi = ((long long)d.b * F + kk) * a.c_total + a.c0 + c;
          float2 v;
          if (a.w_dtype == SETK_C128) {
            const double* p = reinterpret_cast<const double*>(a.w) + 2 * wi;
            v = make_float2((float)p[0], (float)p[1]);
          } else {
            const float* p = reinterpret_cast<const float*>(a.w) + 2 * wi;
            v = make_float2(p[0], p[1]);
          }
          sm.w[c * kAwWPitch + kk] = v;
        }
        named_bar_sync(kAwBarBack, kAwBackThreads);
#pragma unroll
        for (int c = 0; c < C; ++c) { wk[c] = sm.w[c * kAwWPitch + k]; wm[c] = sm.w[c * kAwWPitch + km]; }
      }
      // ---- apply of tile n ----
      const int s = n & (kAwSlots - 1);
      mbar_wait(&sm.z_full[s], (unsigned)(n >> 2) & 1u);
      float2* zt = sm.z + s * kAwJobs * SETK_ZSLOT;
      const bool post = a.post_mask != nullptr;
      if (nt == TT && !post) {
#pragma unroll
        for (int j = 0; j < TT; ++j) item(zt, j, k, km, tw, wk, wm, d.b, d.t0, false);
      } else {
        for (int j = 0; j < nt; ++j) item(zt, j, k, km, tw, wk, wm, d.b, d.t0, post);
      }
      if (warp == (n & 3) && lane < nt) {            // bin 128 (its own mirror), frame `lane`
        float2 w128[C];
#pragma unroll
        for (int c = 0; c < C; ++c) w128[c] = sm.w[c * kAwWPitch + 128];
        item(zt, lane, 128, 128, make_float2(-1.0f, -0.0f), w128, w128, d.b, d.t0, post);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.zi_full[s]);
      // ---- flush of tile n - 1 (its inverse FFT ran while we applied tile n) ----
      if (have_prev) {
        const int sp = (n - 1) & (kAwSlots - 1);
        mbar_wait(&sm.fr_full[sp], (unsigned)((n - 1) >> 2) & 1u);
        flush(sp, prev.t0, (int)(prev.flags & AW_NT), seg_prev, (prev.flags & AW_SEG_BEGIN) != 0);
        if (prev.flags & AW_SEG_END) finish_segment(seg_prev, prev.flags);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.z_empty[sp]);
      }
      prev = d; seg_prev = seg; have_prev = true;
    }
    if (!more) break;
  }
  if (have_prev) {
    const int sp = (n - 1) & (kAwSlots - 1);
    mbar_wait(&sm.fr_full[sp], (unsigned)((n - 1) >> 2) & 1u);
    flush(sp, prev.t0, (int)(prev.flags & AW_NT), seg_prev, (prev.flags & AW_SEG_BEGIN) != 0);
    if (prev.flags & AW_SEG_END) finish_segment(seg_prev, prev.flags);
  }
}

__global__ void __maxnreg__(SETK_AW_LAUNCH_REGS) apply_istft_ws_kernel(ApplyIstftArgs a) {
  SETK_DYN_SMEM(float, smem);
  AwSmem sm;
  sm.carve(smem, a.g.hop);
  const int tid = threadIdx.x;
  const int q = sched_quota(a.sched, gridDim.x);
  const int total = sched_prefix(a.sched, a.sched.B);
  const int lo = blockIdx.x * q;
  const int hi = imin(lo + q, total);
  if (lo >= hi) return;
  const int T_cap = (a.n_out + 2 * a.g.pad + a.g.hop - 1) / a.g.hop;

  for (int n = tid; n < kNfft; n += kAwThreads) {
    const float w = a.window[n];
    sm.win[n] = 0.5f * w;
    sm.wsyn[n] = w * (1.0f / 512.0f);
  }
  for (int n = tid; n < kM; n += kAwThreads) {
    const float s = a.wsq[n] + a.wsq[n + kM];
    sm.rw[n] = s > SETK_TINY32 ? 1.0f / s : 1.0f;
  }
  twiddle_table_fill(sm.twtab, tid, kAwThreads);
  if (tid == 0) {
    mbar_init(&sm.bar_audio[0], 1);
    mbar_init(&sm.bar_audio[1], 1);
    for (int s = 0; s < kAwSlots; ++s) {
      mbar_init(&sm.z_full[s], kAwGroupThreads / 32);
      mbar_init(&sm.zi_full[s], kAwBackThreads / 32);
      mbar_init(&sm.fr_full[s], 1);
      mbar_init(&sm.z_empty[s], kAwBackThreads / 32);
    }
    AwGen g;
    g.x = lo; g.halo_done = 0; g.done = 0; g.count = 0; g.avail = 0;
    *sm.gen = g;
  }
  const bool vec_ok = ((a.N & 3) == 0) && ((a.g.hop & 3) == 0) && ((a.g.pad & 3) == 0) &&
                      ((reinterpret_cast<uintptr_t>(a.audio) & 15) == 0);
  __syncthreads();
  // 64 registers at launch; the inverse-FFT warpgroup hands 16 per thread to the BACK warpgroup
  // (whose threads keep a bin pair's eight complex weights in registers)
  if (tid >= kAwFftThreads + kAwIfftThreads) {
    setmaxnreg_inc<SETK_AW_BACK_REGS>();
    aw_back_role(a, sm, lo, hi, T_cap, vec_ok);
  } else if (tid >= kAwFftThreads) {
    setmaxnreg_dec<SETK_AW_IFFT_REGS>();
    aw_ifft_role(a, sm, lo, hi, T_cap, vec_ok);
  } else {
    aw_fft_role(a, sm, lo, hi, T_cap, vec_ok);
  }
}

int apply_istft_ws_tt() { return kAwTT; }
bool apply_istft_ws_supported(const Geometry& g) {
  return g.n_fft == 512 && g.C == 4 && g.hop == kM;
}

cudaError_t run_apply_istft_ws(const ApplyIstftArgs& a, int n_ctas, void* stream) {
  const size_t smem = AwSmem::bytes(a.g.hop);
  cudaError_t e = cudaFuncSetAttribute(apply_istft_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  return launch(apply_istft_ws_kernel, dim3(n_ctas), dim3(kAwThreads), smem, stream, false, a);
}

}  // namespace setk
