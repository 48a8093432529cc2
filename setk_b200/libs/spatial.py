"""
setk_b200.libs.spatial -- GPU mirror of the reference's spatial features
(scripts/sptk/libs/spatial.py): the same function names, argument names,
defaults and array axes; the arithmetic runs in libsetk_b200.so (csrc/spatial.cu:
setk_ipd, setk_directional_feats, setk_gcc_phat, setk_msc).  No CPU path.

    linear_tdoa_grid   spatial.py:11-34    (host: a 513 x 181 table of constants)
    gcc_phat_linear    spatial.py:37-60
    gcc_phat_diag      spatial.py:63-92
    srp_phat_linear    spatial.py:95-123
    msc                spatial.py:126-160
    ipd                spatial.py:163-181
    directional_feats  spatial.py:184-208

Inputs may be numpy arrays or torch tensors (the same kind comes back).  dtypes
follow the reference for complex64 spectrograms: ipd float32, everything the
reference promotes through a complex128 operand (GCC/SRP, MSC, directional
features with a complex128 steering vector) float64.

Differences from the reference, all deliberate:
  * msc reproduces the code AS WRITTEN, including `np.sum(np.diagonal(icc))`
    without an axis (:153), which adds the grand total of the diagonal terms to
    every cell before the max-normalisation;
  * srp_phat_linear does not compute the (0, 1) pair twice (the reference
    evaluates it once before the pair loop and discards it when N > 2,
    :116-123); for N == 2 it keeps the reference's behaviour of ignoring the
    normalize / apply_floor arguments (:116 passes only **kwargs);
  * only the geometry table (linear_tdoa_grid's omega / tau) is computed on the
    host -- it is a few hundred constants, not data.
"""
import numpy as np
import torch

from .. import plan as _plan
from .utils import EPSILON, default_device  # noqa: F401  (EPSILON re-exported like the reference)

__all__ = ["linear_tdoa_grid", "gcc_phat_linear", "gcc_phat_diag", "srp_phat_linear", "msc", "ipd",
           "directional_feats"]


def _dev(*arrays):
    for a in arrays:
        if torch.is_tensor(a) and a.device.type == "cuda":
            return a.device
    return default_device()


def _back(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t


def _tdoa(dist, speed=343, num_bins=513, samp_doa=True, sample_frequency=16000, num_doa=181,
          max_doa=np.pi):
    """omega (F,), tau (D,) of linear_tdoa_grid (spatial.py:21-33)."""
    dist = np.abs(dist)
    if samp_doa:
        doa_samp = np.linspace(0, max_doa, num_doa)
        tau = np.cos(doa_samp) * dist / speed
    else:
        max_tdoa = dist / speed
        tau = np.linspace(max_tdoa, -max_tdoa, num_doa)
    omega = np.linspace(0, sample_frequency / 2, num_bins) * 2 * np.pi
    return omega, tau


def linear_tdoa_grid(dist, speed=343, num_bins=513, samp_doa=True, sample_frequency=16000,
                     num_doa=181, max_doa=np.pi):
    """Transform matrix T_ij = exp(-j omega_i tau_j) for a linear array, F x D complex128."""
    omega, tau = _tdoa(dist, speed, num_bins, samp_doa, sample_frequency, num_doa, max_doa)
    return np.exp(-1j * np.outer(omega, tau))


def gcc_phat_linear(si, sj, dij, normalize=True, apply_floor=True, **kwargs):
    """GCC-PHAT for a linear array: si, sj T x F, dij the microphone distance -> T x D."""
    was_numpy = not torch.is_tensor(si)
    omega, tau = _tdoa(dij, **kwargs)
    dev = _dev(si, sj)
    out = _plan.gcc_phat(torch.as_tensor(si, device=dev), torch.as_tensor(sj, device=dev), omega, tau,
                         normalize=normalize, apply_floor=apply_floor)
    return _back(out, was_numpy)


def gcc_phat_diag(si, sj, angle_delta, d, speed=343, num_doas=121, sr=16000, normalize=True,
                  num_bins=513, apply_floor=True):
    """GCC-PHAT between diagonal microphones of a circular array of diameter d -> T x D."""
    was_numpy = not torch.is_tensor(si)
    doa_samp = np.linspace(0, np.pi * 2, num_doas)
    tau = np.cos(angle_delta - doa_samp) * d / speed
    omega = np.linspace(0, sr / 2, num_bins) * 2 * np.pi
    dev = _dev(si, sj)
    out = _plan.gcc_phat(torch.as_tensor(si, device=dev), torch.as_tensor(sj, device=dev), omega, tau,
                         normalize=normalize, apply_floor=apply_floor)
    return _back(out, was_numpy)


def srp_phat_linear(S, d, normalize=True, apply_floor=True, **kwargs):
    """SRP-PHAT for a linear array: S N x T x F, d the microphone positions -> T x D."""
    if type(d) is not list and type(d) is not tuple:
        raise ValueError("Now only support linear arrays(in python list/tuple type)")
    N = S.shape[0]
    if N != len(d):
        raise ValueError("{:d} microphones available, while get {:d}-channel STFT".format(len(d), N))
    if S.ndim == 2:
        raise ValueError("Only one-channel STFT available")
    was_numpy = not torch.is_tensor(S)
    dev = _dev(S)
    St = torch.as_tensor(S, device=dev)
    if N == 2:
        omega, tau = _tdoa(d[1] - d[0], **kwargs)
        # spatial.py:116-118: the two-microphone case uses gcc_phat_linear's defaults
        return _back(_plan.gcc_phat(St[0], St[1], omega, tau, normalize=True, apply_floor=True),
                     was_numpy)
    srp = None
    for i in range(N):
        for j in range(i + 1, N):
            omega, tau = _tdoa(d[j] - d[i], **kwargs)
            if srp is None:
                srp = torch.zeros((St.shape[1], len(tau)), dtype=torch.float64, device=dev)
            _plan.gcc_phat(St[i], St[j], omega, tau, normalize=normalize, apply_floor=apply_floor,
                           out=srp)
    return _back(srp * 2 / (N * (N - 1)), was_numpy)


def msc(spectrogram, context=1, normalize=True):
    """MSC (magnitude squared coherence): spectrogram N x T x F -> T x F."""
    was_numpy = not torch.is_tensor(spectrogram)
    out = _plan.msc(torch.as_tensor(spectrogram, device=_dev(spectrogram)), context=context,
                    normalize=normalize)
    return _back(out, was_numpy)


def ipd(si, sj, cos=False, sin=False):
    """IPD / cosIPD / [cosIPD, sinIPD] of two T x F spectrograms."""
    was_numpy = not torch.is_tensor(si)
    dev = _dev(si, sj)
    mode = 0 if not cos else (2 if sin else 1)
    out = _plan.ipd(torch.as_tensor(si, device=dev), torch.as_tensor(sj, device=dev), mode)
    return _back(out, was_numpy)


def directional_feats(spectrogram, steer_vector, df_pair=None):
    """Directional features: spectrogram M x F x T, steer_vector M x F -> T x F."""
    was_numpy = not torch.is_tensor(spectrogram)
    dev = _dev(spectrogram, steer_vector)
    S = torch.as_tensor(spectrogram, device=dev)
    if S.dim() != 3:
        raise ValueError("directional_feats expects an M x F x T spectrogram")
    out = _plan.directional_feats(S[None], torch.as_tensor(steer_vector, device=dev), pairs=df_pair)[0]
    return _back(out, was_numpy)
