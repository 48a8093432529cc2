"""
setk_b200.synth -- seeded synthetic multichannel utterances + oracle IRM masks
for benchmarks and parity tests.  Data generation only:
it runs outside every timed region and is not part of the product path, so it
may use torch ops freely (conv1d, torch.stft).

Per utterance u (seed = base_seed + u):
  target  s[n] : white Gaussian amplitude-modulated by a 4 Hz raised cosine
                 (speech-like time sparsity, energy in every bin)
  image   x_c = h_c * s, h_c a random 64-tap decaying FIR, h_c[0] dominant
  noise   v_c = sum of C+2 independent white sources through their own FIRs
                 + sensor noise at -20 dB re the target image
  mix at 5 dB SNR on channel 0, scaled to max|x| = 0.5, float32 (C, N)
  mask    IRM of channel 0: |S| / sqrt(|S|^2 + |V|^2 + eps)   (the reference's
          compute_mask.py:85-87,107 formula), float32 (T, F)
"""
import math

import torch

EPS32 = 1.1920928955078125e-07


def _fir_bank(gen, n_filters, taps, device):
    decay = torch.exp(-torch.arange(taps, device=device, dtype=torch.float32) / 8.0)
    h = torch.randn((n_filters, taps), generator=gen, device=device) * decay
    lead = h.abs().amax(dim=1) + 0.5
    h[:, 0] = torch.where(h[:, 0] >= 0, lead, -lead)
    return h


def _filter(x, h):
    """x (S, N), h (S, C, taps) -> (S, C, N) causal FIR per (source, channel)."""
    S, C, taps = h.shape
    N = x.shape[-1]
    xp = torch.nn.functional.pad(x[:, None, :], (taps - 1, 0))          # (S,1,N+taps-1)
    w = h.flip(-1).reshape(S * C, 1, taps)
    y = torch.nn.functional.conv1d(xp.reshape(1, S, -1), w, groups=S)   # (1, S*C, N)
    return y.reshape(S, C, N)


def make_utterance(C, N, seed, device, sr=16000, snr_db=5.0, taps=64):
    """Returns (mix (C,N) f32, target image (C,N), noise image (C,N))."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    t = torch.arange(N, device=device, dtype=torch.float32) / sr
    phase = torch.rand((), generator=gen, device=device) * 2 * math.pi
    env = 0.55 + 0.45 * torch.cos(2 * math.pi * 4.0 * t + phase)
    s = torch.randn((1, N), generator=gen, device=device) * env
    nsrc = C + 2
    v = torch.randn((nsrc, N), generator=gen, device=device)
    h_s = _fir_bank(gen, C, taps, device).reshape(1, C, taps)
    h_v = _fir_bank(gen, nsrc * C, taps, device).reshape(nsrc, C, taps)
    tgt = _filter(s, h_s)[0]                                            # (C,N)
    noise = _filter(v, h_v).sum(dim=0)                                  # (C,N)
    noise = noise + 0.1 * tgt.std() * torch.randn((C, N), generator=gen, device=device)
    g = torch.sqrt(tgt[0].pow(2).mean() / (noise[0].pow(2).mean() * 10**(snr_db / 10)))
    noise = noise * g
    mix = tgt + noise
    scale = 0.5 / mix.abs().max()
    return (mix * scale).float(), (tgt * scale).float(), (noise * scale).float()


def irm_mask(tgt0, noise0, frame_len=512, frame_hop=256, n_fft=512, center=True):
    """IRM (T,F) float32 from channel-0 target / noise images (torch.stft)."""
    win = torch.hann_window(frame_len, periodic=True, device=tgt0.device)
    kw = dict(n_fft=n_fft, hop_length=frame_hop, win_length=frame_len, window=win,
              center=center, pad_mode="reflect", return_complex=True)
    S = torch.stft(tgt0, **kw).abs()
    V = torch.stft(noise0, **kw).abs()
    irm = S / torch.sqrt(S * S + V * V + EPS32)
    return irm.transpose(-1, -2).contiguous().float()


def make_batch(B, C, N, base_seed=20240923, device="cuda", frame_len=512, frame_hop=256,
               n_fft=512, center=True, first=0):
    """Returns audio (B,C,N) f32 and IRM masks (B,T,F) f32 on `device`."""
    device = torch.device(device)
    audio = torch.empty((B, C, N), dtype=torch.float32, device=device)
    masks = None
    for u in range(B):
        mix, tgt, noise = make_utterance(C, N, base_seed + first + u, device)
        audio[u] = mix
        m = irm_mask(tgt[0], noise[0], frame_len, frame_hop, n_fft, center)
        if masks is None:
            masks = torch.empty((B,) + tuple(m.shape), dtype=torch.float32, device=device)
        masks[u] = m
    return audio, masks
