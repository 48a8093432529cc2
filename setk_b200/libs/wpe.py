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
  * facted_wpd (wpe.py:113-177) is a composition of this library's kernels: one WPE step with
    an external variance (setk_wpe_step), the 2-class CGMM (setk_cgmm_stft), two mask-weighted
    covariances (setk_cov: weights 1/lambda and the speech mask), the fp64 MVDR solve
    (setk_weights) and setk_apply.  Its per-bin output phase follows this library's eigenvector
    convention (the reference's is LAPACK's, SURVEY.md finding 4).
"""
import numpy as np
import torch

from .. import plan as _plan
from .utils import default_device, get_logger

logger = get_logger(__name__)

__all__ = ["wpe", "facted_wpd"]


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


def facted_wpd(obs, cgmm_iters=10, wpd_iters=3, taps=10, delay=3, context=1, update_alpha=False):
    """
    Joint dereverberation & denoising, factored WPD (wpe.py:113-177).
        obs  N x T x F complex (or B x N x T x F)
    returns (tf_mask T x F x 2 float32, wpd_enh T x F complex64), with a leading batch axis if given.
    """
    from .. import _lib
    was_numpy = isinstance(obs, np.ndarray)
    x = torch.as_tensor(obs)
    if not torch.is_complex(x) or x.dim() not in (3, 4):
        raise RuntimeError("facted_wpd expects a complex spectrogram, N x T x F")
    batched = x.dim() == 4
    if not batched:
        x = x[None]
    B, N, T, F = x.shape
    logger.info(f"Facted WPD: F = {F}, N = {N}, T = {T}")
    dev = x.device if x.device.type == "cuda" else default_device()
    stft = x.to(dev).to(torch.complex64).permute(0, 1, 3, 2).contiguous()          # B x N x F x T
    enh, masks = None, None

    def check(status, what):
        bad = status.nonzero().flatten().tolist()
        if bad:
            raise np.linalg.LinAlgError(f"Singular matrix in facted_wpd/{what} (batch entries {bad})")

    for i in range(wpd_iters):
        logger.info(f"Facted WPD: iter = {i + 1}/{wpd_iters}...")
        der, inv_lam, st = _plan.wpe_step(stft, enh, taps=taps, delay=delay, context=context)
        check(st, "wpe")
        masks, st = _plan.cgmm_from_stft(der, 2, cgmm_iters, update_alpha=update_alpha)   # B x 2 x T x F
        check(st, "cgmm")
        Rd = _plan.covariance(der, inv_lam)              # sum_t der der^H / lambda (any scale: MVDR)
        Rs = _plan.covariance(der, masks[:, 0].contiguous())
        w, st, _ = _plan.weights(_lib.BF_MVDR, Rs, Rn=Rd)
        check(st, "mvdr")
        enh = _plan.apply_weights(der, w)                # B x F x T
    tf_mask = masks.permute(0, 2, 3, 1)                  # B x T x F x 2
    out = enh.transpose(1, 2)                            # B x T x F
    if not batched:
        tf_mask, out = tf_mask[0], out[0]
    if was_numpy:
        return tf_mask.cpu().numpy(), out.cpu().numpy()
    return tf_mask, out
