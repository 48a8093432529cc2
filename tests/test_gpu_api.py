"""GPU tier: the reference-facing Python API (setk_b200.libs) on the real library."""
import numpy as np
import pytest
import torch

from oracle import beamformer_oracle as bo
from oracle import stft_oracle as so

pytestmark = pytest.mark.gpu


def test_libs_numpy_in_numpy_out(cuda):
    from setk_b200.libs import beamformer as BF
    from setk_b200.libs import stft as ST
    from setk_b200.libs import utils as U
    U.set_default_device(None)
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((5, 20000)) * 0.1).astype(np.float32)
    kw = dict(frame_len=512, frame_hop=256, center=True, transpose=False)
    S = np.stack([ST.forward_stft(x[c], **kw) for c in range(5)])
    So = so.multichannel_stft(x, round_power_of_two=True, **kw)
    assert isinstance(S, np.ndarray) and bo.rel_inf(S, So) <= 2e-5
    T, F = S.shape[2], S.shape[1]
    mask = rng.uniform(0, 1, (T, F)).astype(np.float32)
    for bf, kind in ((BF.MvdrBeamformer(F), "mvdr"), (BF.GevdBeamformer(F), "gevd"),
                     (BF.PmwfBeamformer(F), "pmwf")):
        enh = bf.run(mask, S)
        ref = bo.run_supervised(kind, mask.astype(np.float64), So)
        assert bo.rel_inf(bo.align_phase(enh, ref)[0], ref) <= 1e-4
        y = U.inverse_stft(enh, norm=0.5, **kw)
        yo = so.inverse_stft(bo.align_phase(ref, enh)[0], norm=0.5, **kw)
        assert y.shape == yo.shape and bo.rel_inf(y, yo) <= 1e-4
    # torch in -> torch out, on the device
    St = U.forward_stft(torch.from_numpy(x[0]).cuda(), **kw)
    assert St.is_cuda and St.dtype == torch.complex64


def test_no_cpu_fallback(cuda):
    """CPU tensors are not silently processed by the CUDA library."""
    from setk_b200 import plan as P
    with pytest.raises(Exception):
        pl = P.StftPlan(1, 512, 256, True, True, "hann", 1, 4000, torch.device("cpu"))
        pl.stft(torch.zeros(1, 1, 4000))
