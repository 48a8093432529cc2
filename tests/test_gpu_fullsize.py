"""
GPU tier, BASELINE.json's full sizes (config 2: 4 ch x 160000 samples, 512/256,
batch 256): size-independent properties the domain offers, since the oracle
cannot be run on the whole batch in seconds.

  * identity: a weight vector that selects channel 0 makes the fused
    apply+iSTFT reproduce channel 0 (STFT -> iSTFT round trip at full length);
  * linearity of apply+iSTFT in the weights;
  * covariance structure: Hermitian, non-negative diagonal,
    Rs * sum(m) + Rn * sum(1-m) = sum x x^H  (checked through the all-ones mask);
  * MVDR is distortionless on every bin of every utterance: w^H d = 1;
  * fused and explicit-STFT routes agree on sampled utterances;
  * determinism: the same launch twice is bit-identical;
  * the oracle itself on a sample of utterances of the batch (<= 1e-4).
"""
import numpy as np
import pytest
import torch

import parity_cases as pc
from oracle import beamformer_oracle as bo

pytestmark = pytest.mark.gpu
B, C, N = 256, 4, 160000


@pytest.fixture(scope="module")
def batch():
    from setk_b200 import synth
    dev = torch.device("cuda:0")
    a, m = synth.make_batch(8, C, N, device=dev)
    reps = B // 8
    # distinct gains per copy so that utterances differ
    gains = torch.linspace(0.5, 1.0, reps, device=dev).repeat_interleave(8)[:, None, None]
    return (a.repeat(reps, 1, 1) * gains).contiguous(), m.repeat(reps, 1, 1).contiguous()


def test_identity_round_trip_full_length(cuda, batch):
    from setk_b200 import plan as P
    audio, _ = batch
    pl = P.StftPlan(C, 512, 256, True, True, "hann", B, N, cuda)
    w = torch.zeros((B, 257, C), dtype=torch.complex64, device=cuda)
    w[:, :, 0] = 1.0
    y = pl.apply_istft(audio, w)
    assert y.shape == (B, 160000)                       # 256 * (626 - 1): bit-exact bookkeeping
    err = (y - audio[:, 0, :160000]).abs().max() / audio[:, 0].abs().max()
    assert float(err) <= 2e-6
    # linearity in the weights
    w2 = torch.zeros_like(w)
    w2[:, :, 1] = 0.5j
    y2 = pl.apply_istft(audio, w2)
    y12 = pl.apply_istft(audio, w + w2)
    assert float((y12 - (y + y2)).abs().max() / y12.abs().max()) <= 2e-6
    pl.close()


def test_covariance_properties_full_batch(cuda, batch):
    from setk_b200 import plan as P
    audio, mask = batch
    pl = P.StftPlan(C, 512, 256, True, True, "hann", B, N, cuda)
    Rs, Rn, mx = pl.stft_cov(audio, mask)
    Rs2, Rn2, _ = pl.stft_cov(audio, mask)
    assert torch.equal(Rs, Rs2) and torch.equal(Rn, Rn2)                   # deterministic
    assert torch.equal(Rs, Rs.transpose(-1, -2).conj())                    # exactly Hermitian
    d = torch.diagonal(Rs, dim1=-2, dim2=-1)
    assert float(d.imag.abs().max()) == 0.0 and float(d.real.min()) >= 0.0
    assert torch.equal(mx, audio.abs().amax(dim=(1, 2)))                   # max|x| bit-exact
    # sum_t m x x^H + sum_t (1-m) x x^H == sum_t x x^H
    ones = torch.ones_like(mask)
    Ry, _, _ = pl.stft_cov(audio, ones, want_maxabs=False)
    sm = mask.sum(dim=1)[:, :, None, None]                                  # (B,F,1,1)
    T = mask.shape[1]
    lhs = Rs * torch.clamp(sm, min=1e-6) + Rn * torch.clamp(T - sm, min=1e-6)
    rel = (lhs - Ry * T).abs().amax() / (Ry * T).abs().amax()
    assert float(rel) <= 5e-6
    # fused vs explicit-STFT route on a few utterances
    idx = [0, 100, 255]
    S = pl.stft(audio[idx].contiguous()) if len(idx) <= pl.max_batch else None
    Rg = P.covariance(S, mask[idx].contiguous())
    assert float((Rg - Rs[idx]).abs().amax() / Rs[idx].abs().amax()) <= 5e-6
    pl.close()


def test_mvdr_distortionless_and_oracle_sample(cuda, batch):
    from setk_b200 import _lib, plan as P
    from setk_b200.engine import BeamformPipeline
    audio, mask = batch
    pipe = BeamformPipeline(C, "mvdr", max_batch=B, max_samples=N, device=cuda)
    wave, status = pipe.run(audio, mask)
    wave2, _ = pipe.run(audio, mask)
    assert int(status.abs().sum()) == 0
    assert torch.equal(wave, wave2)                                        # deterministic end to end
    assert wave.shape == (B, 160000)
    # peak normalisation: max|y| == max|x| (utils.py:166-168) up to eps
    peak_in = audio.abs().amax(dim=(1, 2))
    assert float(((wave.abs().amax(dim=1) - peak_in).abs() / peak_in).max()) <= 1e-5
    Rs, Rn, _ = pipe.covariances(audio, mask)
    w = pipe.solve(Rs, Rn)[0].to(torch.complex128)
    d = P.weights(_lib.BF_PEVD, Rs.to(torch.complex128))[0]
    wd = (w.conj() * d).sum(-1)
    assert float((wd.abs() - 1).abs().max()) <= 1e-4                       # w^H d = 1 on every bin
    # the oracle on two utterances of the batch
    for b in (3, 200):
        err = pc.mvdr_end_to_end(cuda, audio[b:b + 1].cpu().numpy(), mask[b:b + 1].cpu().numpy())
        assert err <= pc.TOL_E2E, err
