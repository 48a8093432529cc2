"""
setk_b200.libs.utils -- drop-in for the reference's scripts/sptk/libs/utils.py
on the CUDA kernels of libsetk_b200.

Same names, argument names and defaults as the reference:
  nextpow2 (utils.py:25-27), cmat_abs (30-42), write_wav (45-62), read_wav
  (65-92), forward_stft (96-138), inverse_stft (142-173), filekey (208-221),
  get_logger (224-245), EPSILON, MAX_INT16.

Arrays: numpy in -> numpy out (the CLI edge), torch tensor in -> torch tensor
out on the same device.  All arithmetic of forward_stft / inverse_stft runs in
libsetk_b200 kernels (there is no CPU path: without the library or without a
GPU these functions raise).  wav IO uses scipy.io.wavfile because soundfile is
not part of this image; the PCM-16 conversions match soundfile's
(int16/32768 on read, floor(y*32768) on write -- SURVEY.md finding 3) and run
on the GPU as well.
"""
import logging
import os
import warnings

import numpy as np
import torch

from .. import plan as _plan
from ..plan import EPSILON, nextpow2  # noqa: F401  (re-exported)

MAX_INT16 = np.iinfo(np.int16).max
default_format_str = "%(asctime)s [%(pathname)s:%(lineno)s - %(levelname)s ] %(message)s"

__all__ = [
    "forward_stft", "inverse_stft", "get_logger", "filekey", "write_wav",
    "read_wav", "cmat_abs", "nextpow2", "EPSILON", "default_device"
]

_default_device = None


def default_device():
    """Device used for numpy inputs: cuda (required); tests may override."""
    global _default_device
    if _default_device is None:
        if not torch.cuda.is_available():
            raise RuntimeError("setk_b200 needs a CUDA device (there is no CPU fallback)")
        _default_device = torch.device("cuda", torch.cuda.current_device())
    return _default_device


def set_default_device(device):
    global _default_device
    _default_device = None if device is None else torch.device(device)


def _to_tensor(x, dtype=None):
    """-> (tensor on a compute device, was_numpy)."""
    if isinstance(x, torch.Tensor):
        return (x if dtype is None else x.to(dtype)), False
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(default_device()), True


def _back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


# ------------------------------------------------------------ plan cache ----
_plans = {}


def _window_key(window):
    if isinstance(window, str):
        return window
    return ("array", np.asarray(window, dtype=np.float64).tobytes())


def get_plan(num_channels, frame_len, frame_hop, center, round_power_of_two, window, nsamps,
             device, batch=1):
    """Cached StftPlan; re-created when a longer signal or larger batch arrives."""
    key = (num_channels, frame_len, frame_hop, bool(center), bool(round_power_of_two),
           _window_key(window), str(device))
    pl = _plans.get(key)
    if pl is None or pl.max_samples < nsamps or pl.max_batch < batch:
        if pl is not None:
            pl.close()
        cap = max(nsamps, 0 if pl is None else pl.max_samples)
        bcap = max(batch, 1 if pl is None else pl.max_batch)
        pl = _plan.StftPlan(num_channels, frame_len=frame_len, frame_hop=frame_hop, center=center,
                            round_power_of_two=round_power_of_two, window=window,
                            max_batch=bcap, max_samples=cap, device=device)
        _plans[key] = pl
    return pl


# ---------------------------------------------------------------- helpers ---
def cmat_abs(cmat):
    """utils.py:30-42"""
    if isinstance(cmat, torch.Tensor):
        if not cmat.is_complex():
            raise RuntimeError(
                "function cmat_abs expect complex as input, but got {}".format(cmat.dtype))
        return torch.sqrt(cmat.real**2 + cmat.imag**2)
    if not np.iscomplexobj(cmat):
        raise RuntimeError(
            "function cmat_abs expect complex as input, but got {}".format(cmat.dtype))
    return np.sqrt(cmat.real**2 + cmat.imag**2)


def write_wav(fname, samps, sr=16000, normalize=True):
    """utils.py:45-62 (soundfile's default PCM_16 subtype)."""
    import scipy.io.wavfile as wavfile
    if isinstance(samps, torch.Tensor):
        samps = samps.detach().cpu().numpy()
    samps = np.asarray(samps)
    if samps.dtype != np.int16:          # int16 = PCM-16 already (setk_apply_istft_pcm16): written as is
        samps = samps.astype("float32" if normalize else "int16")
    if samps.ndim != 1 and samps.shape[0] < samps.shape[1]:
        samps = np.transpose(samps)
        samps = np.squeeze(samps)
    fdir = os.path.dirname(str(fname))
    if fdir and not os.path.exists(fdir):
        os.makedirs(fdir)
    if samps.dtype != np.int16:
        pcm = np.clip(np.floor(samps.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
    else:
        pcm = samps
    wavfile.write(str(fname), sr, pcm)


def read_wav(fname, beg=0, end=None, normalize=True, sr=16000, raw_pcm16=False):
    """utils.py:65-92: returns C x N (or N) float32.  raw_pcm16=True returns the int16 samples of a
    PCM-16 file untouched (the int16/32768 conversion then happens on the device)."""
    import scipy.io.wavfile as wavfile
    ret_sr, data = wavfile.read(fname)
    if sr != ret_sr:
        raise RuntimeError(f"Expect sr={sr} of {fname}, get {ret_sr} instead")
    data = data[beg:end]
    if raw_pcm16 and data.dtype == np.int16:
        return np.ascontiguousarray(np.transpose(data)) if data.ndim != 1 else data
    if data.dtype == np.int16:
        samps = data.astype(np.float32) / np.float32(32768.0) if normalize else data.astype(
            np.float32)
    elif data.dtype == np.int32:
        samps = (data.astype(np.float64) / 2147483648.0).astype(np.float32) if normalize \
            else (data >> 16).astype(np.float32)
    elif data.dtype == np.uint8:
        samps = ((data.astype(np.float32) - 128.0) / 128.0) if normalize else (
            (data.astype(np.float32) - 128.0) * 256.0)
    else:  # float wav
        samps = data.astype(np.float32) if normalize else (data * 32768.0).astype(np.float32)
    if samps.ndim != 1:
        samps = np.transpose(samps)
    return samps


# return F x T or T x F (tranpose=True)
def forward_stft(samps,
                 frame_len=1024,
                 frame_hop=256,
                 round_power_of_two=True,
                 center=False,
                 window="hann",
                 apply_abs=False,
                 apply_log=False,
                 apply_pow=False,
                 transpose=True):
    """
    STFT wrapper (utils.py:96-138), computed by setk_stft.
    samps: mono vector, numpy or torch.  Returns complex64 (or float32 with
    apply_abs/pow/log), F x T or T x F (transpose=True).
    """
    if apply_log and not apply_abs:
        warnings.warn("Ignore apply_abs=False because apply_log=True")
        apply_abs = True
    if samps.ndim != 1:
        raise RuntimeError("Invalid shape, librosa.stft accepts mono input")
    x, was_np = _to_tensor(samps, torch.float32)
    pl = get_plan(1, frame_len, frame_hop, center, round_power_of_two, window, x.shape[0],
                  x.device)
    stft_mat = pl.stft(x.reshape(1, 1, -1))[0, 0]          # F x T
    if apply_abs:
        stft_mat = cmat_abs(stft_mat)
    if apply_pow:
        stft_mat = stft_mat**2
    if apply_log:
        stft_mat = torch.log(torch.clamp(stft_mat, min=EPSILON))
    if transpose:
        stft_mat = stft_mat.transpose(0, 1)
    return _back(stft_mat, was_np)


# accept F x T or T x F (tranpose=True)
def inverse_stft(stft_mat,
                 frame_len=1024,
                 frame_hop=256,
                 center=False,
                 window="hann",
                 transpose=True,
                 norm=None,
                 power=None,
                 nsamps=None):
    """
    iSTFT wrapper (utils.py:142-173), computed by setk_istft (irfft, window,
    overlap-add, window-sum-square normalisation, trim and -- with `norm` --
    the peak rescale samps * norm / (max|samps| + EPSILON)).
    """
    S, was_np = _to_tensor(stft_mat)
    if not S.is_complex():
        raise RuntimeError("inverse_stft expects a complex STFT matrix")
    if transpose:
        S = S.transpose(0, 1)
    F, T = S.shape
    n_fft = 2 * (F - 1)
    if frame_len > n_fft:
        raise ValueError(f"frame_len {frame_len} does not fit n_fft {n_fft}")
    # a plan whose n_fft is the one implied by the matrix (librosa.istft semantics)
    pl = get_plan(1, frame_len, frame_hop, center, n_fft != frame_len, window,
                  max(frame_hop * T + n_fft, 1), S.device)
    if pl.n_fft != n_fft:
        raise ValueError(f"STFT with {F} bins does not match frame_len {frame_len}")
    nrm = None
    if norm:
        nrm = torch.tensor([float(norm)], dtype=torch.float32, device=S.device)
    samps = pl.istft(S.to(torch.complex64).reshape(1, F, T), n_out=nsamps, norm=nrm)[0]
    if power:
        samps_pow = torch.linalg.vector_norm(samps, 2)**2 / samps.numel()
        samps = samps * torch.sqrt(power / samps_pow)
    return _back(samps, was_np)


def filekey(path):
    """utils.py:208-221"""
    fname = os.path.basename(path)
    if not fname:
        raise ValueError(f"{path}: is directory path?")
    token = fname.split(".")
    if len(token) == 1:
        return token[0]
    else:
        return '.'.join(token[:-1])


def get_logger(name, format_str=default_format_str, date_format="%Y-%m-%d %H:%M:%S", file=False):
    """utils.py:224-245"""

    def get_handler(handler):
        handler.setLevel(logging.INFO)
        handler.setFormatter(logging.Formatter(fmt=format_str, datefmt=date_format))
        return handler

    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    if not logger.handlers:
        if file:
            logger.addHandler(get_handler(logging.FileHandler(name)))
        logger.addHandler(get_handler(logging.StreamHandler()))
    return logger


def check_doa(geometry, doa, online=False):
    """Check value of the DoA (utils.py:248-263)."""
    doas = doa if online else [doa]
    for doa in doas:
        if doa < 0:
            return False
        if geometry == "linear" and doa > 180:
            return False
        if geometry == "circular" and doa >= 360:
            return False
    return True
