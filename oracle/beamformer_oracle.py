"""
oracle/beamformer_oracle.py -- CPU restatement of the reference's mask-based
adaptive beamformer maths (numpy/scipy, float64 by default).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Not imported by the product.

Follows scripts/sptk/libs/beamformer.py:
  do_ban            14-28      solve_pevd        31-63
  rank1_constraint  66-84      compute_covar     87-103
  Beamformer.beamform 220-234  SupervisedBeamformer.run 270-283
  Online* 286-320, 685-728     MVDR 527-539   MPDR 555-590
  PMWF 620-659                 GEVD 674-682
and scripts/sptk/apply_adaptive_beamformer.py:50-71 (VAD mask), 138-177
(mask preparation, post-mask, iSTFT with norm).

Axis conventions are the reference's: obs (N,F,T), mask (T,F), R (F,N,N),
weight (F,N), enhanced (F,T).

Third-party arithmetic outside the reference tree that is *called*, not
restated: LAPACK through numpy (`linalg.eigh`, `linalg.solve`) and scipy
(`linalg.eigh(a, b)`, `linalg.eig`), exactly the entry points the reference
uses (beamformer.py:45,53,57,536).  Their eigenvector sign/phase is
implementation defined (SURVEY.md finding 4); `align_phase` below is the
per-bin alignment every parity comparison applies.
"""
import numpy as np
import scipy.linalg

from .stft_oracle import EPSILON, cmat_abs, inverse_stft


# ----------------------------------------------------------------------------
# free functions
# ----------------------------------------------------------------------------
def _solve_vec(A, b):
    """np.linalg.solve with numpy<2 'stack of vectors' semantics."""
    return np.linalg.solve(A, b[..., None])[..., 0]


def compute_covar(obs, tf_mask):
    """beamformer.py:87-103.  obs (N,F,T), tf_mask (T,F) -> (F,N,N)."""
    x = np.transpose(obs, (1, 0, 2))                       # F N T
    m = np.transpose(tf_mask)[:, None, :]                  # F 1 T
    den = np.maximum(np.sum(m, axis=-1, keepdims=True), 1e-6)
    return np.einsum("fdt,fet->fde", m * x, x.conj()) / den


def solve_pevd(Rs, Rn=None):
    """beamformer.py:31-63.  Principal (generalised) eigenvector, (F,N)."""
    if Rn is None:
        _, vecs = np.linalg.eigh(Rs)
        return vecs[:, :, -1]
    F, N, _ = Rs.shape
    pvec = np.zeros((F, N), dtype=np.complex128)
    for f in range(F):
        try:
            _, vecs = scipy.linalg.eigh(Rs[f], Rn[f])
            pvec[f] = vecs[:, -1]
        except np.linalg.LinAlgError:
            try:
                vals, vecs = scipy.linalg.eig(Rs[f], Rn[f])
                pvec[f] = vecs[:, np.argmax(vals)]
            except np.linalg.LinAlgError:
                raise RuntimeError(
                    f"LinAlgError when computing eig on frequency {f}")
    return pvec


def do_ban(weight, Rn):
    """beamformer.py:14-28.  Blind analytical normalisation."""
    num = np.einsum("...a,...ab,...bc,...c->...", np.conj(weight), Rn, Rn,
                    weight)
    den = np.einsum("...a,...ab,...b->...", np.conj(weight), Rn, weight)
    g = np.sqrt(cmat_abs(num)) / np.maximum(np.real(den), EPSILON)
    return g[:, None] * weight


def rank1_constraint(Rs, Rn=None):
    """beamformer.py:66-84."""
    pvecs = solve_pevd(Rs, Rn=Rn)
    if Rn is not None:
        pvecs = np.einsum("...ab,...b->...a", Rn, pvecs)
    r1 = np.einsum("...a,...b->...ab", pvecs, pvecs.conj())
    scale = np.trace(Rs, axis1=-1, axis2=-2) / np.maximum(
        np.trace(r1, axis1=-1, axis2=-2), EPSILON)
    return scale[..., None, None] * r1


def beamform(weight, obs):
    """beamformer.py:220-234.  weight (F,N), obs (N,F,T) -> (F,T)."""
    if weight.shape[0] != obs.shape[1] or weight.shape[1] != obs.shape[0]:
        raise ValueError("Input obs do not match with weight, " +
                         f"{weight.shape} vs {obs.shape}")
    return np.einsum("fn,nft->ft", weight.conj(), obs)


def compute_covar_mat(num_bins, target_mask, obs):
    """beamformer.py:246-262 (shape checks + compute_covar)."""
    if target_mask.ndim != 2 or target_mask.shape[1] != num_bins:
        raise ValueError("Input mask matrix should be shape as " +
                         f"[num_frames x num_bins], now is {target_mask.shape}")
    if obs.shape[1] != target_mask.shape[1] or obs.shape[
            2] != target_mask.shape[0]:
        raise ValueError("Shape of input obs do not match with " +
                         f"mask matrix, {obs.shape} vs {target_mask.shape}")
    return compute_covar(obs, target_mask)


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------
def mvdr_weight(Rs, Rn):
    """beamformer.py:527-539."""
    d = solve_pevd(Rs)
    num = _solve_vec(Rn, d)
    den = np.einsum("...d,...d->...", d.conj(), num)
    return num / den[..., None]


def mpdr_weight(Rs, Ry, Rn=None):
    """beamformer.py:555-573 (Rn given == whiten)."""
    if Rn is None:
        d = solve_pevd(Rs)
    else:
        d = np.einsum("...ab,...b->...a", Rn, solve_pevd(Rs, Rn))
    num = _solve_vec(Ry, d)
    den = np.einsum("...d,...d->...", d.conj(), num)
    return num / den[..., None]


def gevd_weight(Rs, Rn):
    """beamformer.py:674-682."""
    return solve_pevd(Rs, Rn)


def pmwf_snr(weight, Rs, Rn):
    """beamformer.py:620-630."""
    ps = np.einsum("...fa,...fab,...fb->...", np.conj(weight), Rs, weight)
    pn = np.einsum("...fa,...fab,...fb->...", np.conj(weight), Rn, weight)
    return np.real(ps) / np.maximum(EPSILON, np.real(pn))


def pmwf_weight(Rs, Rn, beta=0, ref_channel=-1, rank1_appro=""):
    """beamformer.py:632-659.  Returns (weight, ref_channel_used)."""
    N = Rs.shape[1]
    if rank1_appro == "eig":
        Rs = rank1_constraint(Rs)
    if rank1_appro == "gev":
        Rs = rank1_constraint(Rs, Rn=Rn)
    G = np.linalg.solve(Rn, Rs)
    den = beta + np.trace(G, axis1=1, axis2=2)
    W = G / den[..., None, None]
    if ref_channel < 0:
        snr = [pmwf_snr(W[..., c], Rs, Rn) for c in range(N)]
        ref = int(np.argmax(snr))
    else:
        ref = ref_channel
    if ref >= N:
        raise RuntimeError("Reference channel ID exceeds total " +
                           f"channels: {ref} vs {N}")
    return W[..., ref], ref


# ----------------------------------------------------------------------------
# SupervisedBeamformer.run and variants
# ----------------------------------------------------------------------------
def run_supervised(kind, mask_s, obs, mask_n=None, ban=False, beta=0,
                   ref_channel=-1, rank1_appro="", return_all=False):
    """
    beamformer.py:270-283 (+ MPDR override 575-590).
    kind in {"mvdr","mpdr","mpdr-whiten","gevd","pmwf"}.
    """
    F = obs.shape[1]
    mask_s = np.asarray(mask_s)
    mn = (1 - mask_s) if mask_n is None else np.asarray(mask_n)
    if kind in ("mpdr", "mpdr-whiten"):
        Rs = compute_covar_mat(F, mask_s, obs)
        Ry = compute_covar_mat(F, np.ones_like(mask_s), obs)
        Rn = compute_covar_mat(F, mn, obs) if (kind == "mpdr-whiten"
                                               or ban) else None
        w = mpdr_weight(Rs, Ry, Rn=Rn if kind == "mpdr-whiten" else None)
    else:
        Rn = compute_covar_mat(F, mn, obs)
        Rs = compute_covar_mat(F, mask_s, obs)
        if kind == "mvdr":
            w = mvdr_weight(Rs, Rn)
        elif kind == "gevd":
            w = gevd_weight(Rs, Rn)
        elif kind == "pmwf":
            w, _ = pmwf_weight(Rs, Rn, beta=beta, ref_channel=ref_channel,
                               rank1_appro=rank1_appro)
        else:
            raise ValueError(f"unknown beamformer kind {kind}")
    if ban:
        w = do_ban(w, Rn)
    enh = beamform(w, obs)
    if return_all:
        return enh, w, Rs, Rn
    return enh


class OnlineState(object):
    """beamformer.py:286-320: exponentially forgotten Rs/Rn."""

    def __init__(self, num_bins, num_channels, alpha=0.8):
        self.shape = (num_bins, num_channels, num_channels)
        self.reset_stats(alpha)

    def reset_stats(self, alpha=0.8):
        self.Rs = np.zeros(self.shape, dtype=np.complex128)
        self.Rn = np.zeros(self.shape, dtype=np.complex128)
        self.alpha = alpha
        self.reset = True

    def run(self, kind, mask_s, obs, mask_n=None, ban=False):
        F = obs.shape[1]
        mn = (1 - mask_s) if mask_n is None else mask_n
        Rn = compute_covar_mat(F, mn, obs)
        Rs = compute_covar_mat(F, mask_s, obs)
        # NOTE: the reference never clears self.reset (beamformer.py:314), so
        # phi stays 1 for every chunk; kept as written.
        phi = 1 if self.reset else (1 - self.alpha)
        self.Rs = self.Rs * self.alpha + phi * Rs
        self.Rn = self.Rn * self.alpha + phi * Rn
        w = mvdr_weight(self.Rs, self.Rn) if kind == "mvdr" else gevd_weight(
            self.Rs, self.Rn)
        return beamform(do_ban(w, Rn) if ban else w, obs)


# ----------------------------------------------------------------------------
# CLI-level pieces (apply_adaptive_beamformer.py)
# ----------------------------------------------------------------------------
def compute_vad_masks(spectrogram, proportion):
    """apply_adaptive_beamformer.py:50-71.  spectrogram F x T -> (T x F bool, index)."""
    energy_mat = cmat_abs(spectrogram)
    energy_vec = np.sort(energy_mat.flatten())
    filter_energy = np.sum(energy_vec) * (1 - proportion)
    csum = np.cumsum(energy_vec)
    over = np.nonzero(csum > filter_energy)[0]
    if over.size:
        index = int(over[0])
        threshold = energy_vec[index]
    else:
        index = energy_vec.shape[0]
        threshold = energy_vec[-1]
    return (energy_mat < threshold).transpose(), index


def prepare_masks(speech_mask, interf_mask, stft_mat, vad_proportion=1):
    """apply_adaptive_beamformer.py:138-158."""
    if interf_mask is None:
        speech_mask = np.minimum(speech_mask, 1)
    F = stft_mat.shape[1]
    if speech_mask.shape[0] == F and speech_mask.shape[1] != F:
        speech_mask = np.transpose(speech_mask)
        if interf_mask is not None:
            interf_mask = np.transpose(interf_mask)
    if 0.5 < vad_proportion < 1:
        vad_mask, _ = compute_vad_masks(stft_mat[0], vad_proportion)
        speech_mask = np.where(vad_mask, 1.0e-4, speech_mask)
        if interf_mask is not None:
            interf_mask = np.where(vad_mask, 1.0e-4, interf_mask)
    return speech_mask, interf_mask


def enhance_utterance(samps, speech_mask, kind="mvdr", interf_mask=None,
                      ban=False, post_mask=False, vad_proportion=1,
                      beta=0, ref_channel=-1, rank1_appro="",
                      frame_len=512, frame_hop=256, window="hann",
                      center=True, round_power_of_two=True,
                      stft_dtype=np.complex128):
    """
    One pass of the per-utterance loop of apply_adaptive_beamformer.py:130-177:
    multichannel STFT -> mask prep -> beamformer.run -> (post-mask) ->
    inverse_stft(norm=max|x|).  samps (C,N) float32.
    Returns (enhanced samples float64, enhanced STFT (F,T), weight (F,N)).
    """
    from .stft_oracle import multichannel_stft
    kw = dict(frame_len=frame_len, frame_hop=frame_hop, window=window,
              center=center, transpose=False)
    obs = multichannel_stft(samps, round_power_of_two=round_power_of_two,
                            out_dtype=stft_dtype, **kw)
    norm = np.max(np.abs(samps))                      # data_handler.py:398-400
    ms, mi = prepare_masks(speech_mask, interf_mask, obs, vad_proportion)
    enh, w, _, _ = run_supervised(kind, ms, obs, mask_n=mi, ban=ban, beta=beta,
                                  ref_channel=ref_channel,
                                  rank1_appro=rank1_appro, return_all=True)
    if post_mask:
        enh = enh * np.transpose(ms)
    y = inverse_stft(enh, norm=norm, **kw)
    return y, enh, w


# ----------------------------------------------------------------------------
# parity helpers
# ----------------------------------------------------------------------------
def align_phase(test, ref, axis=-1):
    """
    Multiply each bin of `test` by the unit-modulus scalar that best aligns it
    with `ref` (least squares): s_f = <ref_f, test_f> / |<ref_f, test_f>|.
    test/ref: (F, K).  Returns aligned copy of test and the scalars.
    """
    inner = np.sum(ref * np.conj(test), axis=axis, keepdims=True)
    mag = np.abs(inner)
    s = np.where(mag > 0, inner / np.maximum(mag, 1e-300), 1.0)
    return test * s, s


def rel_inf(test, ref):
    """global relative infinity-norm error."""
    return float(np.max(np.abs(test - ref)) / max(np.max(np.abs(ref)), 1e-300))
