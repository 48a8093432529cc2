"""
oracle/cgmm_oracle.py -- CPU restatement of the reference's CGMM mask estimator
(the mask producer of BASELINE config 3).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Not imported by the product.

What it restates (scripts/sptk/libs/cluster.py)
  * 94-130   Covariance: Hermitian-symmetrise, eigh, eigenvalues scaled by their
             maximum and floored at EPSILON; R^-1 = V diag(1/w) V^H; log det = sum log w
  * 187-231  CgDistribution.update_parameters / log_pdf
               R_k   = sum_t (gamma_k M / phi_k) y y^H / max(sum_t gamma_k, eps)
               phi_k = max(|y^H R_k^-1 y|, eps) / M
               log N = -M log phi_k - log det R_k
  * 234-287  Cgmm.update / predict (posterior over classes with priors alpha,
             log-sum-exp shifted by the per-TF maximum, denominator floored)
  * 396-465  CgmmTrainer: K = 2 deterministic start (R_0 = sum_t y y^H / T,
             R_1 = I) or a start from given posteriors; `train` = EM iterations
  * estimate_cgmm_masks.py:36-60  masks K x F x T -> K x T x F, class 0 when K = 2,
             float32 on disk
Everything after the (complex64) observations is float64 / complex128, as in the
reference; the one exception there is the K = 2 start, whose sum_t y y^H is a
complex64 einsum (see start_dtype).

Pinned by (oracle/make_golden.py, tests/test_oracle_golden.py), with
start_dtype=complex64 so that the start matches the reference's to the bit:
  * doc/adaptive_beamformer/asset/egs.wav -> the documented command's mask
    (tests/golden/doc_adaptive_beamformer.npz "mask", produced by the reference's
    CgmmTrainer through oracle/ref_shim.py):  max |diff| <= 1e-7 (float32 storage)
  * config-3 mixture (tests/golden/ref_configs.npz "cfg3/mask_cgmm"): same bound.
With the complex128 start the two differ from those fixtures by 1.6e-7 / 7.8e-7 on
average and 1.3e-4 / 1.9e-3 in the worst of 94 576 / 48 222 cells: that is the
reference's own sensitivity to float32 rounding of its start, and the yardstick
for the CUDA path (float32 STFT) against the fixtures.
"""
import numpy as np

EPSILON = np.finfo(np.float32).eps  # utils.py:16


def factor_covariance(R):
    """
    cluster.py:94-130.  R (..., M, M) -> (R^-1, log det) of the *scaled* matrix
    (eigenvalues divided by the largest, floored at EPSILON).
    """
    R = (R + np.conj(np.swapaxes(R, -1, -2))) / 2
    w, v = np.linalg.eigh(R)
    w = w / np.maximum(np.amax(w, axis=-1, keepdims=True), EPSILON)
    w = np.maximum(w, EPSILON)
    R_inv = np.einsum("...xy,...y,...zy->...xz", v, 1 / w, v.conj())
    return R_inv, np.sum(np.log(w), axis=-1)


def quadratic_form(obs, R_inv):
    """phi * M of cluster.py:203-205,442-444: max(|y^H R^-1 y|, eps).  obs F x M x T."""
    q = np.einsum("...xt,...xy,...yt->...t", obs.conj(), R_inv, obs)
    return np.maximum(np.abs(q), EPSILON)


def posterior(phi, log_det, alpha, M):
    """cluster.py:256-283 (Cgmm.predict without Q).  phi K x F x T -> gamma K x F x T."""
    log_pdf = -M * np.log(phi) - log_det[..., None]
    log_pdf = log_pdf - np.amax(log_pdf, 0, keepdims=True)
    nominator = np.exp(log_pdf) * alpha[..., None]
    denominator = np.sum(nominator, 0, keepdims=True)
    return nominator / np.maximum(denominator, EPSILON)


def log_likelihood(phi, log_det, alpha, M):
    """The Q the reference logs (cluster.py:270-274)."""
    log_pdf = -M * np.log(phi) - log_det[..., None]
    return np.mean(np.log(np.sum(np.exp(log_pdf) * alpha[..., None], 0)))


def weighted_covariance(obs, weight, gamma):
    """sum_t weight y y^H / max(sum_t gamma, eps)   (cluster.py:197-201, 437-440)."""
    den = np.maximum(np.sum(gamma, -1), EPSILON)
    R = np.einsum("...t,...xt,...yt->...xy", weight, obs, obs.conj())
    return R / den[..., None, None]


def cgmm_masks(stft, num_classes=2, num_iters=20, init_gamma=None, update_alpha=False,
               return_all=False, start_dtype=np.complex128):
    """
    CgmmTrainer(stft, num_classes, gamma=init_gamma, update_alpha=...).train(num_iters)
    followed by the CLI's transpose (estimate_cgmm_masks.py:52-60).

    stft        M x F x T complex (the reference feeds complex64)
    init_gamma  None, F x T (K = 2: stacked with its complement, cluster.py:428-429)
                or K x F x T
    start_dtype np.complex64 reproduces the reference's K = 2 start to the bit (its
                sum_t y y^H / T is a complex64 einsum, cluster.py:419-420); the default
                accumulates it in complex128 like everything after it
    returns     T x F float64 (K = 2) or K x T x F; with return_all also the
                K x F x T posteriors of every iteration (index 0 = the start)
    """
    obs = np.einsum("mft->fmt", np.asarray(stft)).astype(np.complex128)
    F, M, T = obs.shape
    K = num_classes
    if init_gamma is None:
        if K != 2:
            raise ValueError("a start without posteriors exists for 2 classes only "
                             "(the reference draws np.random.uniform otherwise)")
        o = obs.astype(start_dtype)
        Rs = (np.einsum("...dt,...et->...de", o, o.conj()) / T).astype(np.complex128)
        R = np.stack([Rs, np.broadcast_to(np.eye(M, dtype=np.complex128), (F, M, M))])
    else:
        gamma = np.asarray(init_gamma, dtype=np.float64)
        if K == 2 and gamma.ndim == 2:
            gamma = np.stack([gamma, 1 - gamma])
        R = weighted_covariance(obs, gamma, gamma)
    R_inv, log_det = factor_covariance(R)
    phi = quadratic_form(obs, R_inv) / M
    alpha = np.ones([K, F]) / K
    gamma = posterior(phi, log_det, alpha, M)
    history = [gamma]
    for _ in range(num_iters):
        R = weighted_covariance(obs, gamma * M / phi, gamma)      # M-step
        R_inv, log_det = factor_covariance(R)
        phi = quadratic_form(obs, R_inv) / M
        if update_alpha:
            alpha = np.mean(gamma, -1)
        gamma = posterior(phi, log_det, alpha, M)                 # E-step
        history.append(gamma)
    masks = np.transpose(gamma, (0, 2, 1))
    out = masks[0] if K == 2 else masks
    return (out, history) if return_all else out
