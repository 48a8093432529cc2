"""
oracle/stft_oracle.py -- CPU restatement of the reference's STFT / iSTFT.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Not imported by the product.

What it restates
  * scripts/sptk/libs/utils.py:25-27    nextpow2
  * scripts/sptk/libs/utils.py:30-42    cmat_abs
  * scripts/sptk/libs/utils.py:96-138   forward_stft  (wrapper over librosa.stft)
  * scripts/sptk/libs/utils.py:142-173  inverse_stft  (wrapper over librosa.istft)
and the third-party arithmetic those wrappers call, which is NOT in the
reference tree: librosa==0.8.1 (requirements.txt:2) `core.spectrum.stft`,
`core.spectrum.istft`, `filters.get_window`, `filters.window_sumsquare`,
`util.pad_center`, `util.frame`, `util.fix_length`.  The published algorithm
of that version is restated in `stft()` / `istft()` below:

  stft : w = get_window(window, win_length, fftbins=True)   (periodic), float64
         w zero-padded *centred* to n_fft (left pad (n_fft-win_length)//2)
         center=True  -> y = np.pad(y, n_fft//2, mode="reflect")
         T = 1 + (len(y) - n_fft)//hop ; frame t = y[t*hop : t*hop+n_fft]
         S[:, t] = rfft(w * frame)   in float64, stored complex64 for f32 input
  istft: y[t*hop : t*hop+n_fft] += w * irfft(S[:, t])
         wss[t*hop : ...]       += w**2 ;  y[wss > tiny] /= wss[wss > tiny]
         length None & center  -> y[n_fft//2 : -n_fft//2]
         length given          -> fix_length(y[start:], length)

Pinned by: replay of doc/adaptive_beamformer/asset/egs.wav -> pmwf-0*.wav to
<= 1 LSB of PCM-16 through the reference's own beamformer code with this
module standing in for librosa (oracle/make_golden.py).
"""
import math

import numpy as np
import scipy.signal

EPSILON = np.finfo(np.float32).eps  # utils.py:16


def nextpow2(window_size):
    """utils.py:25-27"""
    return 2**math.ceil(math.log2(window_size))


def cmat_abs(cmat):
    """utils.py:30-42"""
    if not np.iscomplexobj(cmat):
        raise RuntimeError(
            "function cmat_abs expect complex as input, but got {}".format(
                cmat.dtype))
    return np.sqrt(cmat.real**2 + cmat.imag**2)


def make_window(window, frame_len):
    """
    Window of length frame_len as float64 (librosa.filters.get_window with
    fftbins=True; "sqrthann" per utils.py:116-117,156-157).
    """
    if isinstance(window, str):
        if window == "sqrthann":
            return scipy.signal.get_window("hann", frame_len,
                                           fftbins=True)**0.5
        return scipy.signal.get_window(window, frame_len, fftbins=True)
    window = np.asarray(window, dtype=np.float64)
    if window.shape != (frame_len,):
        raise ValueError(
            f"Window size mismatch: {window.shape[0]} != {frame_len}")
    return window


def pad_center(win, n_fft):
    """librosa.util.pad_center for 1-D data."""
    n = win.shape[0]
    if n > n_fft:
        raise ValueError(f"Target size ({n_fft}) must be at least input size ({n})")
    lpad = (n_fft - n) // 2
    return np.pad(win, (lpad, n_fft - n - lpad), mode="constant")


def num_frames(nsamps, n_fft, hop, center):
    """Frame count of librosa.stft: integer bookkeeping, bit-exact."""
    padded = nsamps + 2 * (n_fft // 2) if center else nsamps
    if padded < n_fft:
        raise ValueError(f"Input too short: {nsamps} samples for n_fft={n_fft}")
    return 1 + (padded - n_fft) // hop


def istft_length(num_frames_, n_fft, hop, center, nsamps=None):
    """Output length of librosa.istft (length=nsamps)."""
    if nsamps is not None:
        return nsamps
    full = n_fft + hop * (num_frames_ - 1)
    return full - 2 * (n_fft // 2) if center else full


def stft(y, n_fft, hop, win_length, window="hann", center=True,
         out_dtype=np.complex128):
    """
    librosa 0.8.1 stft (float64 arithmetic).  out_dtype=np.complex64 gives the
    reference's storage dtype for float32 input.
    """
    y = np.asarray(y)
    if y.ndim != 1:
        raise RuntimeError("Invalid shape, librosa.stft accepts mono input")
    w = pad_center(make_window(window, win_length), n_fft)
    yd = y.astype(np.float64)
    if center:
        if y.shape[0] < n_fft // 2 + 1:
            raise ValueError("reflect padding needs more than n_fft//2 samples")
        yd = np.pad(yd, n_fft // 2, mode="reflect")
    T = num_frames(y.shape[0], n_fft, hop, center)
    idx = np.arange(n_fft)[:, None] + hop * np.arange(T)[None, :]
    frames = yd[idx]                                   # n_fft x T
    spec = np.fft.rfft(w[:, None] * frames, axis=0)    # F x T, complex128
    return spec.astype(out_dtype)


def istft(S, hop, win_length, window="hann", center=True, length=None,
          out_dtype=np.float64):
    """librosa 0.8.1 istft (float64 arithmetic)."""
    S = np.asarray(S)
    n_fft = 2 * (S.shape[0] - 1)
    w = pad_center(make_window(window, win_length), n_fft)
    if length is None:
        T = S.shape[1]
    else:
        padded_length = length + n_fft if center else length
        T = min(S.shape[1], int(np.ceil(padded_length / hop)))
    expected = n_fft + hop * (T - 1)
    y = np.zeros(expected, dtype=np.float64)
    wss = np.zeros(expected, dtype=np.float64)
    frames = np.fft.irfft(S[:, :T].astype(np.complex128), n=n_fft, axis=0)
    frames *= w[:, None]
    wsq = w**2
    for t in range(T):
        y[t * hop:t * hop + n_fft] += frames[:, t]
        wss[t * hop:t * hop + n_fft] += wsq
    # librosa compares against tiny of the *output* dtype (float32 for c64 in)
    tiny = np.finfo(np.float32).tiny if S.dtype == np.complex64 else np.finfo(
        np.float64).tiny
    nz = wss > tiny
    y[nz] /= wss[nz]
    if length is None:
        if center:
            y = y[n_fft // 2:-(n_fft // 2)]
    else:
        start = n_fft // 2 if center else 0
        y = y[start:]
        if y.shape[0] > length:
            y = y[:length]
        elif y.shape[0] < length:
            y = np.pad(y, (0, length - y.shape[0]), mode="constant")
    return y.astype(out_dtype)


def forward_stft(samps,
                 frame_len=1024,
                 frame_hop=256,
                 round_power_of_two=True,
                 center=False,
                 window="hann",
                 apply_abs=False,
                 apply_log=False,
                 apply_pow=False,
                 transpose=True,
                 out_dtype=np.complex128):
    """utils.py:96-138 (same argument names and defaults)."""
    if apply_log and not apply_abs:
        apply_abs = True
    samps = np.asarray(samps)
    if samps.ndim != 1:
        raise RuntimeError("Invalid shape, librosa.stft accepts mono input")
    n_fft = nextpow2(frame_len) if round_power_of_two else frame_len
    stft_mat = stft(samps, n_fft, frame_hop, frame_len, window=window,
                    center=center, out_dtype=out_dtype)
    if apply_abs:
        stft_mat = cmat_abs(stft_mat)
    if apply_pow:
        stft_mat = np.power(stft_mat, 2)
    if apply_log:
        stft_mat = np.log(np.maximum(stft_mat, EPSILON))
    if transpose:
        stft_mat = np.transpose(stft_mat)
    return stft_mat


def inverse_stft(stft_mat,
                 frame_len=1024,
                 frame_hop=256,
                 center=False,
                 window="hann",
                 transpose=True,
                 norm=None,
                 power=None,
                 nsamps=None,
                 out_dtype=np.float64):
    """utils.py:142-173 (same argument names and defaults)."""
    if transpose:
        stft_mat = np.transpose(stft_mat)
    samps = istft(stft_mat, frame_hop, frame_len, window=window, center=center,
                  length=nsamps, out_dtype=out_dtype)
    if norm:
        samps_norm = np.linalg.norm(samps, np.inf)
        samps = samps * norm / (samps_norm + EPSILON)
    if power:
        samps_pow = np.linalg.norm(samps, 2)**2 / samps.size
        samps = samps * np.sqrt(power / samps_pow)
    return samps


def multichannel_stft(samps, **stft_kwargs):
    """data_handler.py:492-503: per-channel forward_stft, stacked C x F x T."""
    samps = np.asarray(samps)
    if samps.ndim == 1:
        return forward_stft(samps, **stft_kwargs)
    return np.stack(
        [forward_stft(samps[c], **stft_kwargs) for c in range(samps.shape[0])])


def pcm16_from_float(y):
    """
    soundfile/libsndfile float -> PCM_16 as the shipped doc vectors imply
    (SURVEY.md finding 3): floor(y * 32768) clipped to int16.
    utils.py:45-62 (write_wav -> sf.write default subtype).
    """
    v = np.floor(np.asarray(y, dtype=np.float64) * 32768.0)
    return np.clip(v, -32768, 32767).astype(np.int16)


def float_from_pcm16(x):
    """soundfile read dtype=float32: int16 / 32768 (utils.py:80-92)."""
    return (np.asarray(x, dtype=np.float32) / np.float32(32768.0)).astype(
        np.float32)
