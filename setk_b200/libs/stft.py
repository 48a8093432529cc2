"""
setk_b200.libs.stft -- the STFT entry points under the module name the
north-star spec uses (`sptk.libs.stft`).  In the reference they live in
scripts/sptk/libs/utils.py:96-173; this module re-exports the CUDA-backed
implementations together with the batched plan object.
"""
from ..plan import StftPlan, make_window, nextpow2  # noqa: F401
from .utils import forward_stft, inverse_stft  # noqa: F401

__all__ = ["forward_stft", "inverse_stft", "StftPlan", "make_window", "nextpow2"]
