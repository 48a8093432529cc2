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
