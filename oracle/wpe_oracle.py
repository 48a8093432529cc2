"""
oracle/wpe_oracle.py -- CPU restatement of the reference's WPE dereverberation
(the pre-processor of BASELINE config 4).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Not imported by the product.

What it restates (scripts/sptk/libs/wpe.py)
  * 14-30   compute_tap_mat: the delayed stack  yt[k N + n, t] = x[n, t - (k + delay)]
  * 33-56   compute_lambda: per-frame power averaged over channels, box-smoothed over
            +-context frames (edge frames divide by the number of frames present),
            floored at EPSILON
  * 59-79   wpe_step:  R = sum_t yt yt^H / lambda,  r = sum_t yt x^H / lambda,
            G = solve(R, r),  z = x - G^H yt
  * 82-110  wpe: num_iters steps, lambda from the previous estimate (first: the input)
The reference runs this in the dtype of its input: complex64 observations give
float32 einsum accumulations and a single-precision LAPACK solve.  `dtype`
selects that (np.complex64, to pin against the reference's own output) or
complex128 (what the CUDA kernels compute in, checked against this).

Pinned by (oracle/make_golden.py "wpe", tests/test_oracle_golden.py): small mixtures
run through the REFERENCE's wpe() under oracle/ref_shim.py (tests/golden/ref_wpe.npz);
in complex64 mode the restatement reproduces them to float32 rounding.
"""
import numpy as np

EPSILON = np.finfo(np.float32).eps  # utils.py:16


def tap_matrix(obs, taps, delay):
    """wpe.py:14-30.  obs F x N x T -> F x (N taps) x T."""
    F, N, T = obs.shape
    y = np.zeros([F, N * taps, T], dtype=obs.dtype)
    for k in range(taps):
        d = k + delay
        if d >= T:
            break
        y[:, k * N:(k + 1) * N, d:] = obs[:, :, :T - d]
    return y


def frame_variance(z, ctx=0):
    """wpe.py:33-56.  z F x N x T -> lambda F x T."""
    L = np.mean(z.real**2 + z.imag**2, axis=1)
    T = L.shape[1]
    lam = np.zeros_like(L)
    cnt = np.zeros(T)
    for c in range(-ctx, ctx + 1):
        s, e = max(c, 0), min(T, T + c)
        lam[:, s:e] += L[:, max(-c, 0):min(T, T - c)]
        cnt[s:e] += 1
    return np.maximum(lam / cnt, EPSILON)


def wpe(reverb, taps=10, delay=3, context=1, num_iters=3, dtype=np.complex128, return_filters=False):
    """
    wpe.py:82-110.  reverb F x N x T complex -> dereverberated F x N x T (dtype).
    """
    x = np.asarray(reverb).astype(dtype)
    yt = tap_matrix(x, taps, delay)
    z = x
    G = None
    for _ in range(num_iters):
        lam = frame_variance(z, ctx=context)
        yn = yt / lam[:, None, :]
        R = np.einsum("...mt,...nt->...mn", yn, yt.conj())
        r = np.einsum("...mt,...nt->...mn", yn, x.conj())
        G = np.linalg.solve(R, r)
        z = x - np.einsum("...na,...nb->...ab", G.conj(), yt)
    return (z, G) if return_filters else z


def wpe_step(x, yt, lam):
    """wpe.py:58-77 for a given variance lam (F x T): one weighted least-squares prediction step."""
    yn = yt / lam[:, None, :]
    R = np.einsum("...mt,...nt->...mn", yn, yt.conj())
    r = np.einsum("...mt,...nt->...mn", yn, x.conj())
    G = np.linalg.solve(R, r)
    return x - np.einsum("...na,...nb->...ab", G.conj(), yt)


def facted_wpd(obs, cgmm_iters=10, wpd_iters=3, taps=10, delay=3, context=1, update_alpha=False,
               dtype=np.complex128):
    """
    wpe.py:113-177 (factored WPD: joint dereverberation and denoising).  obs N x T x F complex.
    Per iteration: one WPE step (variance = compute_lambda of the observations, then |previous
    output|^2, floored at EPSILON), a 2-class CGMM on the dereverberated channels, Rd = sum_t
    der der^H / lambda / T, Rs = compute_covar(der, mask), sv = principal eigenvector of Rs,
    w = Rd^-1 sv / (sv^H Rd^-1 sv), output = w^H der.
    Returns (tf_mask T x F x 2, wpd_enh T x F) like the reference.  `dtype` = the precision of the
    dereverberation stage (np.complex64 follows the reference's dtype flow for complex64 input).
    """
    from oracle import beamformer_oracle as bo
    from oracle import cgmm_oracle as co
    x = np.einsum("ntf->fnt", np.asarray(obs)).astype(dtype)             # F x N x T
    yt = tap_matrix(x, taps, delay)
    enh, gamma = None, None
    for i in range(wpd_iters):
        lam = frame_variance(x, ctx=context) if i == 0 else np.abs(enh) ** 2
        lam = np.maximum(lam, EPSILON)
        der = wpe_step(x, yt, lam)                                        # F x N x T
        der_r = np.einsum("fnt->nft", der)
        _, hist = co.cgmm_masks(der_r, 2, cgmm_iters, update_alpha=update_alpha, return_all=True,
                                start_dtype=np.complex64 if dtype == np.complex64 else np.complex128)
        gamma = hist[-1]                                                  # K x F x T
        Rd = np.einsum("...nt,...mt->...nm", der / lam[:, None], der.conj()) / der.shape[-1]
        Rs = bo.compute_covar(der_r, gamma[0].T)
        sv = bo.solve_pevd(Rs)
        Rd_inv_sv = np.linalg.solve(Rd, sv[..., None])[..., 0]
        den = np.einsum("...d,...d->...", sv.conj(), Rd_inv_sv)
        weight = Rd_inv_sv / den[:, None]
        enh = np.einsum("...n,...nt->...t", weight.conj(), der)           # F x T
    return gamma.T, enh.T
