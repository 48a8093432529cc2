"""
setk_b200.libs.wpe -- GPU mirror of the reference's WPE dereverberation
(scripts/sptk/libs/wpe.py:82-110 `wpe`, with compute_tap_mat / compute_lambda /
wpe_step 14-79 behind it).

    dereverb = wpe(reverb, taps=10, delay=3, context=1, num_iters=3)   # F x N x T

Same name, arguments, defaults and axes as the reference; the arithmetic runs in
libsetk_b200.so (setk_wpe_stft: one CTA per frequency bin, fp64 normal equations,
LU with partial pivoting; csrc/wpe.cu).  No CPU path.

Differences from the reference:
  * reverb may be a torch tensor or a numpy array (the same kind comes back), with
    an optional leading batch axis (B x F x N x T);
  * the arithmetic is fp64 whatever the input precision (the reference computes
    in complex64 for complex64 input); the result is returned as complex64;
  * an exactly singular normal matrix raises numpy.linalg.LinAlgError, like
    np.linalg.solve in the reference (its CLI catches it per utterance);
  * facted_wpd (wpe.py:113-175) and the WPD beamformer are outside SURVEY.md §8.
"""
import numpy as np
import torch

from .. import plan as _plan
from .utils import default_device, get_logger

logger = get_logger(__name__)

__all__ = ["wpe"]


def wpe(reverb, taps=10, delay=3, context=1, num_iters=3):
    """
    GWPE dereverberation (wpe.py:82-110).
        reverb  complex spectrogram, F x N x T (or B x F x N x T)
    returns the dereverberated spectrogram of the same shape (complex64)
    """
    was_numpy = isinstance(reverb, np.ndarray)
    x = torch.as_tensor(reverb)
    if not torch.is_complex(x) or x.dim() not in (3, 4):
        raise RuntimeError("wpe expects a complex spectrogram, F x N x T")
    batched = x.dim() == 4
    if not batched:
        x = x[None]
    B, F, N, T = x.shape
    logger.info(f"WPE: F = {F}, N = {N}, T = {T}")
    dev = x.device if x.device.type == "cuda" else default_device()
    stft = x.to(dev).to(torch.complex64).permute(0, 2, 1, 3).contiguous()       # B x N x F x T
    out, status = _plan.wpe_from_stft(stft, taps=taps, delay=delay, context=context,
                                      num_iters=num_iters)
    bad = status.nonzero().flatten().tolist()
    if bad:
        raise np.linalg.LinAlgError(f"Singular matrix in wpe (batch entries {bad})")
    out = out.permute(0, 2, 1, 3)                                                # B x F x N x T
    if not batched:
        out = out[0]
    return out.cpu().numpy() if was_numpy else out
