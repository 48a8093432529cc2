"""
oracle/spatial_oracle.py -- CPU restatement of the reference's spatial features
(scripts/sptk/libs/spatial.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Not imported by the product.

What it restates
  * 11-34    linear_tdoa_grid: omega = 2 pi linspace(0, sr/2, F), tau from the DOA / TDOA grid
  * 37-60    gcc_phat_linear;  63-92 gcc_phat_diag:  Re(exp(j(angle si - angle sj)) @ exp(-j omega tau)),
             / max(max|.|, EPSILON), floor at 0
  * 95-123   srp_phat_linear: mean over microphone pairs of the per-pair (normalised, floored) GCC
  * 126-160  msc, AS WRITTEN: the diagonal of icc enters as its grand total (np.sum, no axis, :153)
  * 163-181  ipd;  184-208 directional_feats
dtypes are left to numpy exactly like the reference: complex64 spectrograms give float32
phases; the GCC transform and MSC are complex128 / float64.

Pinned by (oracle/make_golden.py "spatial", tests/test_oracle_golden.py): seeded cases run
through the REFERENCE's spatial.py under oracle/ref_shim.py (tests/golden/ref_spatial.npz);
the restatement reproduces every one of them bit for bit (same numpy calls).
"""
import numpy as np

EPSILON = np.finfo(np.float32).eps  # utils.py:16


def tdoa_grid(dist, speed=343, num_bins=513, samp_doa=True, sample_frequency=16000, num_doa=181,
              max_doa=np.pi):
    """spatial.py:21-33 -> (omega (F,), tau (D,))."""
    dist = np.abs(dist)
    if samp_doa:
        tau = np.cos(np.linspace(0, max_doa, num_doa)) * dist / speed
    else:
        max_tdoa = dist / speed
        tau = np.linspace(max_tdoa, -max_tdoa, num_doa)
    omega = np.linspace(0, sample_frequency / 2, num_bins) * 2 * np.pi
    return omega, tau


def gcc_phat(si, sj, omega, tau, normalize=True, apply_floor=True):
    """spatial.py:47-60 / 82-92 for a given grid."""
    coherence = np.exp(1j * (np.angle(si) - np.angle(sj)))
    transform = np.exp(-1j * np.outer(omega, tau))
    spectrum = np.real(coherence @ transform)
    if normalize:
        spectrum = spectrum / np.max(np.maximum(np.abs(spectrum), EPSILON))
    if apply_floor:
        spectrum = np.maximum(spectrum, 0)
    return spectrum


def gcc_phat_linear(si, sj, dij, normalize=True, apply_floor=True, **kwargs):
    omega, tau = tdoa_grid(dij, **kwargs)
    return gcc_phat(si, sj, omega, tau, normalize, apply_floor)


def gcc_phat_diag(si, sj, angle_delta, d, speed=343, num_doas=121, sr=16000, normalize=True,
                  num_bins=513, apply_floor=True):
    tau = np.cos(angle_delta - np.linspace(0, np.pi * 2, num_doas)) * d / speed
    omega = np.linspace(0, sr / 2, num_bins) * 2 * np.pi
    return gcc_phat(si, sj, omega, tau, normalize, apply_floor)


def srp_phat_linear(S, d, normalize=True, apply_floor=True, **kwargs):
    """spatial.py:95-123."""
    N = S.shape[0]
    if N == 2:
        # the reference calls gcc_phat_linear(S[0], S[1], d[1]-d[0], **kwargs): defaults apply
        return gcc_phat_linear(S[0], S[1], d[1] - d[0], **kwargs)
    srp = None
    for i in range(N):
        for j in range(i + 1, N):
            g = gcc_phat_linear(S[i], S[j], d[j] - d[i], normalize, apply_floor, **kwargs)
            srp = g if srp is None else srp + g
    return srp * 2 / (N * (N - 1))


def msc(spectrogram, context=1, normalize=True):
    """spatial.py:126-160 as written."""
    N, T, F = spectrogram.shape
    K = context * 2 + 1
    Y = np.zeros([K, N, T, F], dtype=np.complex128)
    for t in range(T):
        for i, c in enumerate(range(-context, context + 1)):
            Y[i, :, t, :] = spectrogram[:, min(max(t + c, 0), T - 1), :]
    numerator = np.einsum("ab...,bc...->ac...", np.swapaxes(Y, 0, 1), np.conj(Y)) / K
    dig = np.abs(np.diagonal(numerator, axis1=0, axis2=1))
    dig = np.transpose(dig, [2, 0, 1])
    denumerator = np.sqrt(np.einsum("a...,b...->ab...", dig, dig))
    icc = np.abs(numerator / denumerator)
    coh = np.sum(np.diagonal(icc, axis1=0, axis2=1))          # a SCALAR (no axis)
    coh = coh + np.sum(np.sum(icc, axis=0), axis=0)
    coh = coh / (N * (N - 1))
    if normalize:
        coh = coh / np.max(np.abs(coh))
    return coh


def ipd(si, sj, cos=False, sin=False):
    """spatial.py:163-181."""
    ipd_mat = np.angle(si) - np.angle(sj)
    if not cos:
        return np.mod(ipd_mat + np.pi, 2 * np.pi) - np.pi
    cos_ipd = np.cos(ipd_mat)
    if not sin:
        return cos_ipd
    return np.concatenate((cos_ipd, np.sin(ipd_mat)), axis=1)


def directional_feats(spectrogram, steer_vector, df_pair=None):
    """spatial.py:184-208."""
    M = spectrogram.shape[0]
    arg_s, arg_t = np.angle(spectrogram), np.angle(steer_vector)
    if df_pair is None:
        df_pair = [(i, j) for i in range(M) for j in range(i + 1, M)]
    df = []
    for i, j in df_pair:
        delta_s = arg_s[i] - arg_s[j]
        delta_t = np.expand_dims(arg_t[i] - arg_t[j], 1)
        df.append(np.cos(delta_s - delta_t))
    return np.transpose(np.average(np.stack(df), axis=0))
