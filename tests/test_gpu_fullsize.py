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

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="no CUDA device")]
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


def test_persistent_schedule_batch_compositions(cuda):
    """
    The fused kernels cut the batch's (utterance, tile) sequence into equal runs per CTA, so where
    an utterance is cut depends on the batch around it.  An utterance's results must not: the
    enhanced audio is bit-identical (every sample is the same arithmetic wherever the cut falls),
    the covariances agree to fp32 summation order, for batches of 1, 3, 37 and 300 utterances,
    ragged lengths from one frame to the full 10 s, and an utterance too short for one frame.
    """
    from setk_b200 import plan as P, synth
    dev = cuda
    a, m = synth.make_batch(4, C, N, device=dev)
    T = m.shape[1]
    for Bx in (1, 3, 37, 300):
        pl = P.StftPlan(C, 512, 256, True, True, "hann", Bx, N, dev)
        idx = torch.arange(Bx, device=dev) % 4
        audio = (a[idx] * torch.linspace(0.3, 1.0, Bx, device=dev)[:, None, None]).contiguous()
        mask = m[idx].contiguous()
        lens = torch.full((Bx,), N, dtype=torch.int32, device=dev)
        if Bx >= 3:
            lens[1] = 700                       # three frames
            lens[2] = 100                       # too short for reflect padding: no frame at all
            for b in range(3, Bx):
                lens[b] = int(N * (0.05 + 0.95 * ((b * 7919) % 101) / 100.0))
        Rs, Rn, mx = pl.stft_cov(audio, mask, n_samples=lens)
        w = torch.zeros((Bx, 257, C), dtype=torch.complex64, device=dev)
        w[:, :, 0] = 0.75
        w[:, :, 2] = 0.25j
        y = pl.apply_istft(audio, w, n_samples=lens)
        assert torch.isfinite(Rs).all() and torch.isfinite(y).all()
        for b in sorted({0, 1, 2, Bx // 2, Bx - 1} & set(range(Bx))):
            nb = int(lens[b])
            one = P.StftPlan(C, 512, 256, True, True, "hann", 1, N, dev)
            Rs1, Rn1, mx1 = one.stft_cov(audio[b:b + 1], mask[b:b + 1], n_samples=lens[b:b + 1])
            y1 = one.apply_istft(audio[b:b + 1], w[b:b + 1], n_samples=lens[b:b + 1])
            one.close()
            assert torch.equal(y[b], y1[0]), (Bx, b)
            scale = float(Rs1.abs().amax()) or 1.0
            assert float((Rs[b] - Rs1[0]).abs().amax()) <= 5e-6 * scale, (Bx, b)
            assert float((Rn[b] - Rn1[0]).abs().amax()) <= 5e-6 * max(float(Rn1.abs().amax()), 1e-30), (Bx, b)
            assert float(mx[b]) == float(mx1[0])
            if nb == 100:                       # no frame: zero covariances, silent output
                assert float(Rs[b].abs().amax()) == 0.0 and float(y[b].abs().amax()) == 0.0
        pl.close()
