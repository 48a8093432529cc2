"""
Shared parity checks: the CUDA path (through the C-ABI via setk_b200.plan)
against the CPU oracle on the same seeded inputs.  The same functions run

  * in the CPU tier against tests/emu/libsetk_b200_emu.so (the kernels' sources
    under the CPU execution model) at tiny sizes, and
  * in the GPU tier against libsetk_b200.so on a B200.

Tolerances (float path; integer bookkeeping is asserted exactly):
  STFT / covariance / iSTFT  fp32 kernels vs float64 oracle : rel-inf <= 2e-5
                             (observed ~3e-7; c64 storage alone is 6e-8)
  weights                    fp64 kernel vs LAPACK           : rel-inf <= 1e-9 after
                             per-bin phase alignment (observed 1e-13)
  end-to-end MVDR wave       north_star's bound              : rel-inf <= 1e-4 after
                             per-bin phase alignment of the weights
"""
import numpy as np
import torch

from oracle import beamformer_oracle as bo
from oracle import stft_oracle as so
from setk_b200 import _lib
from setk_b200 import plan as P

TOL_F32 = 2e-5
TOL_W = 1e-9
TOL_E2E = 1e-4


def rng_audio(rng, B, C, N, scale=0.1):
    return (rng.standard_normal((B, C, N)) * scale).astype(np.float32)


def stft_kwargs(frame_len, hop, center, window):
    return dict(frame_len=frame_len, frame_hop=hop, center=center, window=window,
                transpose=False)


def oracle_stft(x, frame_len, hop, center, window, dtype=np.complex128):
    return so.multichannel_stft(x, round_power_of_two=True, out_dtype=dtype,
                                **stft_kwargs(frame_len, hop, center, window))


def check_bookkeeping(device, frame_len, hop, center, lengths):
    """Frame counts / iSTFT lengths are integer work: bit-exact vs the oracle."""
    pl = P.StftPlan(1, frame_len, hop, center, True, "hann", 1, max(lengths), device)
    n_fft = so.nextpow2(frame_len)
    for n in lengths:
        T = so.num_frames(n, n_fft, hop, center)
        assert pl.num_frames(n) == T
        assert pl.istft_length(T) == so.istft_length(T, n_fft, hop, center)
    pl.close()


def check_stft(device, rng, B, C, N, frame_len=512, hop=256, center=True, window="hann",
               n_samples=None):
    x = rng_audio(rng, B, C, N)
    pl = P.StftPlan(C, frame_len, hop, center, True, window, B, N, device)
    S = pl.stft(torch.from_numpy(x).to(device), n_samples=n_samples).cpu().numpy()
    assert S.dtype == np.complex64
    for b in range(B):
        nb = N if n_samples is None else int(n_samples[b])
        if nb < (frame_len if not center else pl.n_fft // 2 + 1):
            assert not S[b].any()        # too short for one frame (librosa raises): zeros
            continue
        So = oracle_stft(x[b, :, :nb], frame_len, hop, center, window)
        Tb = So.shape[-1]
        assert S.shape[1:3] == So.shape[0:2]
        err = bo.rel_inf(S[b, :, :, :Tb], So)
        assert err <= TOL_F32, f"stft rel-inf {err}"
        assert np.all(S[b, :, :, Tb:] == 0)
    pl.close()
    return err


def check_stft_cov(device, rng, B, C, N, frame_len=512, hop=256, center=True, window="hann",
                   with_mask_n=False, n_samples=None, clip=False, mask_ft=False):
    x = rng_audio(rng, B, C, N)
    pl = P.StftPlan(C, frame_len, hop, center, True, window, B, N, device)
    T, F = pl.num_frames(N), pl.num_bins
    hi = 1.3 if clip else 1.0
    ms = rng.uniform(0, hi, (B, T, F)).astype(np.float32)
    mn = rng.uniform(0, 1, (B, T, F)).astype(np.float32) if with_mask_n else None
    tms = torch.from_numpy(ms.transpose(0, 2, 1).copy() if mask_ft else ms).to(device)
    tmn = None if mn is None else torch.from_numpy(
        mn.transpose(0, 2, 1).copy() if mask_ft else mn).to(device)
    Rs, Rn, mx = pl.stft_cov(torch.from_numpy(x).to(device), tms, tmn, n_samples=n_samples,
                             clip_mask=clip, mask_ft=mask_ft)
    Rs, Rn, mx = Rs.cpu().numpy(), Rn.cpu().numpy(), mx.cpu().numpy()
    worst = 0.0
    for b in range(B):
        nb = N if n_samples is None else int(n_samples[b])
        if nb < (frame_len if not center else pl.n_fft // 2 + 1) or (center and nb + pl.n_fft < pl.n_fft):
            # too short for one frame (librosa raises): the batch entry gets zero covariances
            assert not Rs[b].any() and not Rn[b].any()
            assert mx[b] == np.float32(np.max(np.abs(x[b, :, :nb])))
            continue
        So = oracle_stft(x[b, :, :nb], frame_len, hop, center, window)
        Tb = So.shape[-1]
        m_s = ms[b, :Tb].astype(np.float64)
        if clip:
            m_s = np.minimum(m_s, 1)
        m_n = (1 - m_s) if mn is None else mn[b, :Tb].astype(np.float64)
        Rs_o = bo.compute_covar(So, m_s)
        Rn_o = bo.compute_covar(So, m_n)
        es, en = bo.rel_inf(Rs[b], Rs_o), bo.rel_inf(Rn[b], Rn_o)
        assert es <= TOL_F32 and en <= TOL_F32, f"covariance rel-inf {es} {en}"
        # exactly Hermitian, real diagonal (the reference's only asserted property,
        # test/test-beamformer.cc:45)
        assert np.array_equal(Rs[b], np.conj(np.swapaxes(Rs[b], -1, -2)))
        assert mx[b] == np.float32(np.max(np.abs(x[b, :, :nb])))
        worst = max(worst, es, en)
    pl.close()
    return worst


def check_non_power_of_two(device, rng, frame_len=400, hop=160, C=3, N=4000):
    """
    --round-power-of-two false (libs/opts.py:41, utils.py:115): n_fft = frame_len.  Bookkeeping
    bit-exact, STFT / iSTFT round trip and the whole PMWF chain against the oracle.
    """
    from setk_b200.engine import BeamformPipeline
    kw = dict(frame_len=frame_len, frame_hop=hop, center=True, window="hann", transpose=False)
    pl = P.StftPlan(C, frame_len, hop, True, False, "hann", 2, N, device)
    assert pl.n_fft == frame_len and pl.num_bins == frame_len // 2 + 1
    for n in (N, N - 37, 2 * frame_len + 1):
        T = so.num_frames(n, frame_len, hop, True)
        assert pl.num_frames(n) == T
        assert pl.istft_length(T) == so.istft_length(T, frame_len, hop, True)
    x = rng_audio(rng, 2, C, N)
    S = pl.stft(torch.from_numpy(x).to(device)).cpu().numpy()
    for b in range(2):
        So = so.multichannel_stft(x[b], round_power_of_two=False, **kw)
        assert S[b].shape == So.shape
        assert bo.rel_inf(S[b], So) <= TOL_F32
    y = pl.istft(torch.from_numpy(S[:, 0]).to(device)).cpu().numpy()       # channel 0 round trip
    n_rt = y.shape[1]
    assert bo.rel_inf(y[:, frame_len:n_rt - frame_len], x[:, 0, frame_len:n_rt - frame_len]) <= TOL_F32
    pl.close()
    pipe = BeamformPipeline(C, "pmwf-0", frame_len=frame_len, frame_hop=hop, round_power_of_two=False,
                            max_batch=2, max_samples=N, device=device)
    T, F = pipe.plan.num_frames(N), pipe.plan.num_bins
    m = rng.uniform(0.05, 0.95, (2, T, F)).astype(np.float32)
    xs = structured_audio(rng, 2, C, N)
    wave, st = pipe.run(torch.from_numpy(xs).to(device), torch.from_numpy(m).to(device))
    assert int(st.abs().sum()) == 0
    for b in range(2):
        yo, _, _ = bo.enhance_utterance(xs[b], m[b], kind="pmwf", beta=0, frame_len=frame_len,
                                        frame_hop=hop, round_power_of_two=False)
        assert bo.rel_inf(wave[b].cpu().numpy(), yo) <= TOL_E2E
    pipe.plan.close()


def check_cov_generic(device, rng, B, C, F, T):
    X = (rng.standard_normal((B, C, F, T)) + 1j * rng.standard_normal((B, C, F, T))).astype(
        np.complex64)
    m = rng.uniform(0, 1, (B, T, F)).astype(np.float32)
    R = P.covariance(torch.from_numpy(X).to(device), torch.from_numpy(m).to(device)).cpu().numpy()
    for b in range(B):
        Ro = bo.compute_covar(X[b].astype(np.complex128), m[b].astype(np.float64))
        assert bo.rel_inf(R[b], Ro) <= TOL_F32


def random_hpd(rng, F, C, cond=50.0):
    A = rng.standard_normal((F, C, C + 4)) + 1j * rng.standard_normal((F, C, C + 4))
    R = A @ np.conj(np.swapaxes(A, -1, -2)) / (C + 4)
    return R + np.eye(C)[None] * (np.trace(R, axis1=-1, axis2=-2).real[:, None, None] / C / cond)


def random_rank1_plus(rng, F, C):
    """Rs with a clearly dominant eigenvector in every bin (eigen-gap >= ~5)."""
    d = rng.standard_normal((F, C)) + 1j * rng.standard_normal((F, C))
    return 6.0 * d[:, :, None] * np.conj(d[:, None, :]) + 0.3 * random_hpd(rng, F, C)


def check_weights(device, rng, B, F, C, dtype=np.complex128):
    """Every beamformer kind against the oracle (= the reference's LAPACK calls)."""
    Rs = np.stack([random_rank1_plus(rng, F, C) for _ in range(B)]).astype(dtype)
    Rn = np.stack([random_hpd(rng, F, C) for _ in range(B)]).astype(dtype)
    Ry = (Rs + Rn).astype(dtype)
    tRs, tRn, tRy = (torch.from_numpy(a).to(device) for a in (Rs, Rn, Ry))
    tol = TOL_W if dtype == np.complex128 else 5e-4
    Rs64, Rn64, Ry64 = Rs.astype(np.complex128), Rn.astype(np.complex128), Ry.astype(np.complex128)

    def cmp(w, wo, align=True):
        w = w.cpu().numpy().astype(np.complex128)
        worst = 0.0
        for b in range(B):
            wa = bo.align_phase(w[b], wo[b])[0] if align else w[b]
            worst = max(worst, bo.rel_inf(wa, wo[b]))
        assert worst <= tol, f"weights rel-inf {worst}"

    out = {}
    w, st, _ = P.weights(_lib.BF_PEVD, tRs)
    cmp(w, [bo.solve_pevd(Rs64[b]) for b in range(B)])
    wn = w.cpu().numpy()
    assert np.allclose(np.linalg.norm(wn, axis=-1), 1, atol=1e-5)
    assert np.all(np.abs(wn[..., 0].imag) <= 1e-7) and np.all(wn[..., 0].real >= 0)
    w, st1, _ = P.weights(_lib.BF_PEVD, tRs, tRn)
    cmp(w, [bo.solve_pevd(Rs64[b], Rn64[b]) for b in range(B)])
    w, st2, _ = P.weights(_lib.BF_MVDR, tRs, tRn)
    cmp(w, [bo.mvdr_weight(Rs64[b], Rn64[b]) for b in range(B)])
    out["mvdr"] = w
    w, st3, _ = P.weights(_lib.BF_MVDR, tRs, tRn, ban=True)
    cmp(w, [bo.do_ban(bo.mvdr_weight(Rs64[b], Rn64[b]), Rn64[b]) for b in range(B)])
    w, st4, _ = P.weights(_lib.BF_GEVD, tRs, tRn)
    cmp(w, [bo.gevd_weight(Rs64[b], Rn64[b]) for b in range(B)])
    w, st5, _ = P.weights(_lib.BF_MPDR, tRs, None, tRy)
    cmp(w, [bo.mpdr_weight(Rs64[b], Ry64[b]) for b in range(B)])
    w, st6, _ = P.weights(_lib.BF_MPDR_WHITEN, tRs, tRn, tRy)
    cmp(w, [bo.mpdr_weight(Rs64[b], Ry64[b], Rn=Rn64[b]) for b in range(B)])
    for beta in (0.0, 1.0):
        for r1, r1name in ((_lib.RANK1_NONE, ""), (_lib.RANK1_EIG, "eig"), (_lib.RANK1_GEV, "gev")):
            for ref in (-1, C - 1):
                w, st7, used = P.weights(_lib.BF_PMWF, tRs, tRn, beta=beta, ref_channel=ref, rank1=r1)
                exp = [bo.pmwf_weight(Rs64[b], Rn64[b], beta=beta, ref_channel=ref,
                                      rank1_appro=r1name) for b in range(B)]
                cmp(w, [e[0] for e in exp], align=False)   # PMWF is phase invariant
                assert [int(u) for u in used.cpu()] == [e[1] for e in exp]
                assert int(st7.abs().sum()) == 0
    for st in (st, st1, st2, st3, st4, st5, st6):
        assert int(st.abs().sum()) == 0
    # MVDR is distortionless: w^H d = 1
    d = np.stack([bo.solve_pevd(Rs64[b]) for b in range(B)])
    wd = np.sum(np.conj(out["mvdr"].cpu().numpy().astype(np.complex128)) * d, axis=-1)
    assert np.allclose(np.abs(wd), 1, atol=1e-6 if dtype == np.complex128 else 1e-3)


def check_weights_status(device, C=3):
    """Singular / non-PD inputs set status bits instead of failing the call."""
    F = 3
    Rs = np.tile(np.eye(C, dtype=np.complex128), (1, F, 1, 1))
    Rn = np.zeros((1, F, C, C), dtype=np.complex128)          # exactly singular
    w, st, _ = P.weights(_lib.BF_MVDR, torch.from_numpy(Rs).to(device),
                         torch.from_numpy(Rn).to(device))
    assert int(st[0]) & _lib.ST_SINGULAR
    Rn2 = np.tile(np.diag([1.0, -1.0] + [1.0] * (C - 2)).astype(np.complex128), (1, F, 1, 1))
    w, st, _ = P.weights(_lib.BF_GEVD, torch.from_numpy(Rs).to(device),
                         torch.from_numpy(Rn2).to(device))
    assert int(st[0]) & _lib.ST_NOT_PD
    w, st, _ = P.weights(_lib.BF_PMWF, torch.from_numpy(Rs).to(device),
                         torch.from_numpy(Rs.copy()).to(device), ref_channel=C + 2)
    assert int(st[0]) & _lib.ST_BAD_REF


def check_apply_istft(device, rng, B, C, N, frame_len=512, hop=256, center=True, window="hann",
                      post_mask=False, norm=True, n_out=None, n_samples=None):
    x = rng_audio(rng, B, C, N)
    pl = P.StftPlan(C, frame_len, hop, center, True, window, B, N, device)
    T, F = pl.num_frames(N), pl.num_bins
    w = (rng.standard_normal((B, F, C)) + 1j * rng.standard_normal((B, F, C))).astype(np.complex64)
    pm = rng.uniform(0, 1, (B, T, F)).astype(np.float32)
    nm = np.abs(x).max(axis=(1, 2)).astype(np.float32)
    y = pl.apply_istft(torch.from_numpy(x).to(device), torch.from_numpy(w).to(device),
                       post_mask=torch.from_numpy(pm).to(device) if post_mask else None,
                       n_out=n_out, norm=torch.from_numpy(nm).to(device) if norm else None,
                       n_samples=n_samples).cpu().numpy()
    kw = stft_kwargs(frame_len, hop, center, window)
    worst = 0.0
    for b in range(B):
        nb = N if n_samples is None else int(n_samples[b])
        if nb < (frame_len if not center else pl.n_fft // 2 + 1):
            assert not y[b].any()        # too short for one frame (librosa raises): silence
            continue
        So = oracle_stft(x[b, :, :nb], frame_len, hop, center, window)
        Tb = So.shape[-1]
        enh = bo.beamform(w[b].astype(np.complex128), So)
        if post_mask:
            enh = enh * pm[b, :Tb].T
        if n_samples is not None:
            # a ragged batch: every utterance is its own inverse_stft(nsamps=None) -- natural
            # length, peak over that length only -- followed by zeros up to the batch's n_out
            yn = so.inverse_stft(enh, norm=float(nm[b]) if norm else None, nsamps=None, **kw)
            yo = np.zeros(y.shape[1], dtype=yn.dtype)
            k = min(len(yn), len(yo))
            yo[:k] = yn[:k]
        else:
            yo = so.inverse_stft(enh, norm=float(nm[b]) if norm else None, nsamps=n_out, **kw)
        assert yo.shape[0] == y.shape[1], (yo.shape, y.shape)      # integer bookkeeping
        err = bo.rel_inf(y[b], yo)
        assert err <= TOL_F32, f"apply+istft rel-inf {err}"
        worst = max(worst, err)
    pl.close()
    return worst


def check_generic_chain(device, rng, B, C, N, frame_len, hop, center, window):
    """setk_stft -> setk_apply -> setk_istft (explicit STFT route) vs the oracle."""
    x = rng_audio(rng, B, C, N)
    pl = P.StftPlan(C, frame_len, hop, center, True, window, B, N, device)
    T, F = pl.num_frames(N), pl.num_bins
    w = (rng.standard_normal((B, F, C)) + 1j * rng.standard_normal((B, F, C))).astype(np.complex128)
    S = pl.stft(torch.from_numpy(x).to(device))
    enh = P.apply_weights(S, torch.from_numpy(w).to(device))
    nm = np.abs(x).max(axis=(1, 2)).astype(np.float32)
    y = pl.istft(enh, norm=torch.from_numpy(nm).to(device)).cpu().numpy()
    kw = stft_kwargs(frame_len, hop, center, window)
    for b in range(B):
        So = oracle_stft(x[b], frame_len, hop, center, window)
        yo = so.inverse_stft(bo.beamform(w[b], So), norm=float(nm[b]), **kw)
        assert yo.shape[0] == y.shape[1]
        assert bo.rel_inf(y[b], yo) <= TOL_F32
    pl.close()


def mvdr_end_to_end(device, x, mask, kind="mvdr", frame_len=512, hop=256, center=True,
                    window="hann", **bf_kwargs):
    """
    Whole pipeline on one batch vs the oracle's enhance_utterance, with the
    oracle's weights phase-aligned per bin to ours (the reference's sign is
    LAPACK-defined, SURVEY.md finding 4) before the oracle's own
    beamform + iSTFT.  Returns the worst global rel-inf over the batch.
    """
    from setk_b200.engine import BeamformPipeline
    B, C, N = x.shape
    name = {"pmwf": "pmwf-0"}.get(kind, kind)
    pipe = BeamformPipeline(C, beamformer=name, frame_len=frame_len, frame_hop=hop, center=center,
                            window=window, max_batch=B, max_samples=N, device=device, **bf_kwargs)
    tx, tm = torch.from_numpy(x).to(device), torch.from_numpy(mask).to(device)
    y, status = pipe.run(tx, tm)
    assert int(status.abs().sum()) == 0
    Rs, Rn, _ = pipe.covariances(tx, tm)
    Ry = None
    if kind in ("mpdr", "mpdr-whiten"):       # Ry: the all-ones mask (beamformer.py:575-590)
        Ry = pipe.plan.stft_cov(tx, torch.ones_like(tm), None, want_maxabs=False)[0]
    w_ours = pipe.solve(Rs, Rn, Ry)[0].cpu().numpy().astype(np.complex128)
    y = y.cpu().numpy()
    kw = stft_kwargs(frame_len, hop, center, window)
    worst = 0.0
    for b in range(B):
        So = oracle_stft(x[b], frame_len, hop, center, window)
        ms = np.minimum(mask[b].astype(np.float64), 1)
        okind = "pmwf" if kind.startswith("pmwf") else kind
        _, w_o, _, _ = bo.run_supervised(okind, ms, So, ban=bf_kwargs.get("ban", False),
                                         beta=1 if kind == "pmwf-1" else 0,
                                         ref_channel=bf_kwargs.get("pmwf_ref", -1),
                                         rank1_appro=bf_kwargs.get("rank1_appro", ""),
                                         return_all=True)
        w_al, _ = bo.align_phase(w_o, w_ours[b])
        yo = so.inverse_stft(bo.beamform(w_al, So), norm=float(np.max(np.abs(x[b]))), **kw)
        worst = max(worst, bo.rel_inf(y[b], yo))
    return worst


def check_pcm(device, rng):
    """read_wav / write_wav sample conversions: bit-exact integer work."""
    for n in (4096, 4099):
        pcm = rng.integers(-32768, 32768, size=n, dtype=np.int16)
        w = P.pcm16_to_float(torch.from_numpy(pcm).to(device)).cpu().numpy()
        assert np.array_equal(w, so.float_from_pcm16(pcm))
        y = (rng.standard_normal(n) * 0.5).astype(np.float32)
        y[:4] = [1.5, -1.5, 0.99999, -1.0]
        q = P.float_to_pcm16(torch.from_numpy(y).to(device)).cpu().numpy()
        assert np.array_equal(q, so.pcm16_from_float(y))


def pack_cm_blobs(mats, T, F):
    """Kaldi "CM" bytes of every matrix in a uint8 [B][slot] array (slot = a multiple of 16)."""
    from setk_b200.libs.data_handler import compress_kaldi_cm
    slot = (16 + F * (8 + T) + 15) & ~15
    blobs = np.zeros((len(mats), slot), dtype=np.uint8)
    for b, m in enumerate(mats):
        raw = np.frombuffer(compress_kaldi_cm(m), dtype=np.uint8)
        blobs[b, :raw.size] = raw
    return blobs


def check_cm_masks(device, rng, B=3, T=70, F=257):
    """Kaldi CompressedMatrix masks expanded on the device == the archive reader (which the CPU tier
    pins to the reference's own kaldi_io.uncompress): bit-exact, ragged rows zero-filled, a bad
    header flagged and decoded to zeros."""
    import io
    from setk_b200.libs.data_handler import read_kaldi_matrix
    rows = [T] + [int(rng.integers(1, T)) for _ in range(B - 1)]
    mats = [rng.random((r, F)).astype(np.float32) ** 2 for r in rows]
    mats[-1][:, 3] = 0.25                                 # a constant column: degenerate percentiles
    blobs = pack_cm_blobs(mats, T, F)
    blobs[0, 16 + 8 * F:16 + 8 * F + 4] = (64, 65, 192, 193)    # the segments' edges
    expect = np.zeros((B, T, F), dtype=np.float32)
    for b in range(B):
        size = 16 + F * (8 + rows[b])
        expect[b, :rows[b]] = read_kaldi_matrix(io.BytesIO(b"\0BCM " + blobs[b, :size].tobytes()))
    out, st = P.cm_masks(torch.from_numpy(blobs).to(device), T, F)
    assert not st.cpu().numpy().any()
    assert np.array_equal(out.cpu().numpy(), expect)
    bad = blobs.copy()
    bad[1, 8:12] = np.frombuffer(np.int32(T + 1).tobytes(), dtype=np.uint8)     # rows > T
    out, st = P.cm_masks(torch.from_numpy(bad).to(device), T, F)
    assert list(st.cpu().numpy()) == [0, 1] + [0] * (B - 2)
    o = out.cpu().numpy()
    assert not o[1].any() and np.array_equal(o[0], expect[0])


def check_config_fixture(device, name, tol=TOL_E2E, stft_from_oracle=False):
    """
    tests/golden/ref_configs.npz (oracle/make_golden.py): BASELINE.json configs 3 and 4
    with the mask producer / pre-processor run by the reference's own code.
      cfg3: 8 ch, 1024-pt, reference CGMM mask -> GevdBeamformer.run
      cfg4: 6 ch, 512-pt, reference WPE output -> MvdrBeamformer.run
    Both through the reference-facing API (setk_b200.libs) on explicit STFTs.
    """
    import os
    from setk_b200.libs import beamformer as BF
    from setk_b200.libs import utils as U
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_configs.npz"))
    U.set_default_device(device)
    try:
        if name == "cfg3":
            mix, mask = g["cfg3/mix"], np.minimum(g["cfg3/mask_cgmm"], 1)
            kw = dict(frame_len=1024, frame_hop=256, center=True, window="hann", transpose=False)
            if stft_from_oracle:     # CPU tier: keep the emulated run short
                obs = so.multichannel_stft(mix, round_power_of_two=True, out_dtype=np.complex64, **kw)
            else:
                obs = np.stack([U.forward_stft(mix[c], **kw) for c in range(mix.shape[0])])
            enh = BF.GevdBeamformer(obs.shape[1]).run(mask, obs)
            ref = g["cfg3/enh_gevd"].astype(np.complex128)
            # A near-binary CGMM mask leaves a few bins with fewer noise frames than
            # channels: Rn is singular there even in float64, the reference's answer is
            # whatever LAPACK's fallback returns, ours comes from a diagonally loaded Rn.
            # Compare the bins whose noise covariance is well conditioned.
            Rn = bo.compute_covar(obs.astype(np.complex128), 1 - mask.astype(np.float64))
            ev = np.linalg.eigvalsh(Rn)
            good = ev[:, 0] > 1e-5 * ev[:, -1]
            assert good.sum() >= 0.97 * good.size
            assert rel_inf_aligned(enh[good], ref[good]) <= 10 * tol     # cond(Rn) up to 1e5
            assert np.all(np.isfinite(enh))
            if not stft_from_oracle:
                y = U.inverse_stft(enh, norm=float(np.max(np.abs(mix))), **kw)
                assert y.shape == g["cfg3/y_gevd"].shape
        else:
            obs, mask = g["cfg4/stft_wpe"], g["cfg4/mask"]
            enh = BF.MvdrBeamformer(obs.shape[1]).run(mask, obs)
            assert rel_inf_aligned(enh, g["cfg4/enh_mvdr"].astype(np.complex128)) <= tol
    finally:
        U.set_default_device(None)


def rel_inf_aligned(test, ref):
    return bo.rel_inf(bo.align_phase(np.asarray(test, dtype=np.complex128), ref)[0], ref)


# ---------------------------------------------------------------- CGMM masks ---
TOL_CGMM = 1e-6          # fp64 kernels vs the float64 oracle on the same STFT (observed 3e-8:
                         # the float32 rounding of the stored mask)


def structured_audio(rng, B, C, N):
    """A gated common source through short random filters plus sensor noise."""
    x = (rng.standard_normal((B, C, N)) * 0.05).astype(np.float32)
    for b in range(B):
        s = rng.standard_normal(N) * 0.3 * (np.sin(np.arange(N) / (200.0 + 50 * b)) > 0)
        for c in range(C):
            h = rng.standard_normal(8) * np.exp(-np.arange(8) / 3.0)
            x[b, c] += np.convolve(s, h)[:N].astype(np.float32)
    return x


def check_cgmm(device, rng, B, C, N, frame_len=512, hop=256, K=2, iters=4, with_init=False,
               update_alpha=False, n_samples=None):
    """setk_cgmm_masks vs oracle.cgmm_oracle fed the library's own STFT."""
    from oracle import cgmm_oracle as co
    x = structured_audio(rng, B, C, N)
    pl = P.StftPlan(C, frame_len, hop, True, True, "hann", B, N, device)
    xt = torch.from_numpy(x).to(device)
    X = pl.stft(xt, n_samples=n_samples).cpu().numpy()          # (B, C, F, T)
    T, F = X.shape[-1], X.shape[2]
    init = None
    if with_init:
        g = rng.uniform(0.05, 1.0, size=(B, K, T, F))
        init = (g / g.sum(1, keepdims=True)).astype(np.float32)
    masks, status = pl.cgmm_masks(xt, K, iters, init_gamma=None if init is None else torch.from_numpy(init).to(device),
                                  update_alpha=update_alpha, n_samples=n_samples)
    masks = masks.cpu().numpy()
    assert masks.shape == (B, K, T, F) and masks.dtype == np.float32
    assert int(status.cpu().abs().sum()) == 0
    worst = 0.0
    for b in range(B):
        Tb = T if n_samples is None else pl.num_frames(int(n_samples[b]))
        ig = None if init is None else np.transpose(init[b, :, :Tb].astype(np.float64), (0, 2, 1))
        ref = co.cgmm_masks(X[b][:, :, :Tb], K, iters, init_gamma=ig, update_alpha=update_alpha,
                            return_all=True)[1][-1]              # K x F x T
        ref = np.transpose(ref, (0, 2, 1))
        err = float(np.max(np.abs(masks[b, :, :Tb] - ref)))
        assert err <= TOL_CGMM, f"cgmm masks differ by {err}"
        assert np.all(masks[b, :, Tb:] == 0)
        worst = max(worst, err)
    pl.close()
    return worst


def check_cgmm_fixture(device, name, mean_tol=2e-5, max_tol=5e-3):
    """
    From the fixture's audio to the REFERENCE's masks (tests/golden/ref_cgmm.npz,
    CgmmTrainer run by the reference on its float64 STFT).  The library's STFT is
    float32 arithmetic, and EM amplifies that rounding exactly as it amplifies the
    float32 rounding of the reference's own start (oracle/cgmm_oracle.py header:
    1.9e-3 worst cell, 8e-7 mean on the config-3 fixture) -- hence a mean and a
    worst-cell bound rather than one tight rel-inf.
    """
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cgmm.npz"))
    mix = g[name + "/mix"]
    K, iters, upd = (int(v) for v in g[name + "/cfg"])
    init = g[name + "/init_gamma"]                               # F x T or K x F x T
    if init.ndim == 2:
        init = np.stack([init, 1 - init])                        # cluster.py:428-429
    init = np.ascontiguousarray(np.transpose(init, (0, 2, 1)))[None]   # 1 x K x T x F
    C, N = mix.shape
    pl = P.StftPlan(C, 512, 256, True, True, "hann", 1, N, device)
    masks, _ = pl.cgmm_masks(torch.from_numpy(mix[None]).to(device), K, iters,
                             init_gamma=torch.from_numpy(init).to(device), update_alpha=bool(upd))
    d = np.abs(masks[0].cpu().numpy() - g[name + "/masks"])
    pl.close()
    assert d.mean() <= mean_tol and d.max() <= max_tol, (name, float(d.mean()), float(d.max()))
    return float(d.mean()), float(d.max())


def check_cgmm_documented(device, which):
    """
    The documented mask command (estimate_cgmm_masks.py --num-iters 20) from audio:
    "doc"  egs.wav, 5 ch, 512/256  vs the reference's mask (doc_adaptive_beamformer.npz)
    "cfg3" config-3 mixture, 8 ch, 1024/256 vs ref_configs.npz "cfg3/mask_cgmm".
    Returns (mean, max, fraction of cells off by more than 1e-3).
    """
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if which == "doc":
        g = np.load(os.path.join(gold, "doc_adaptive_beamformer.npz"))
        x, ref, fl = so.float_from_pcm16(g["egs_pcm16"]), g["mask"], 512
    else:
        g = np.load(os.path.join(gold, "ref_configs.npz"))
        x, ref, fl = g["cfg3/mix"], g["cfg3/mask_cgmm"], 1024
    C, N = x.shape
    pl = P.StftPlan(C, fl, 256, True, True, "hann", 1, N, device)
    masks, status = pl.cgmm_masks(torch.from_numpy(np.ascontiguousarray(x[None])).to(device), 2, 20)
    m = masks[0, 0].cpu().numpy()
    pl.close()
    assert int(status.cpu().abs().sum()) == 0
    assert m.shape == ref.shape
    d = np.abs(m - ref)
    return float(d.mean()), float(d.max()), float(np.mean(d > 1e-3))


# ------------------------------------------------------------------- WPE ---
TOL_WPE = 2e-6           # fp64 kernels, complex64 output, vs the float64 oracle on the same STFT


def check_wpe(device, rng, B, C, N, frame_len=512, hop=128, taps=4, delay=2, ctx=1, iters=2,
              return_out=False):
    """setk_wpe_stft vs oracle.wpe_oracle on the oracle's complex64 STFT."""
    from oracle import wpe_oracle as wo
    x = structured_audio(rng, B, C, N)
    S = np.stack([oracle_stft(x[b], frame_len, hop, True, "hann", dtype=np.complex64) for b in range(B)])
    out, status = P.wpe_from_stft(torch.from_numpy(S).to(device), taps, delay, ctx, iters)
    out = out.cpu().numpy()
    assert out.shape == S.shape and out.dtype == np.complex64
    assert int(status.cpu().abs().sum()) == 0
    worst = 0.0
    for b in range(B):
        ref = wo.wpe(np.einsum("nft->fnt", S[b]), taps, delay, ctx, iters)       # F x N x T, complex128
        err = bo.rel_inf(out[b], np.einsum("fnt->nft", ref))
        assert err <= TOL_WPE, f"wpe rel-inf {err}"
        worst = max(worst, err)
    return out if return_out else worst


def check_wpe_fixture(device, name):
    """From the fixture's audio to the REFERENCE's wpe() output (tests/golden/ref_wpe.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_wpe.npz"))
    fl, hop, taps, delay, ctx, iters = (int(v) for v in g[name + "/cfg"])
    mix = g[name + "/mix"]
    S = oracle_stft(mix, fl, hop, True, "hann", dtype=np.complex64)[None]
    out, status = P.wpe_from_stft(torch.from_numpy(S).to(device), taps, delay, ctx, iters)
    assert int(status.cpu().abs().sum()) == 0
    err = bo.rel_inf(out[0].cpu().numpy(), g[name + "/derev"])
    assert err <= 1e-5, (name, err)      # the reference ran in complex64; both sit ~1e-8 from float64
    return err


# ---------------------------------------------------------------------------
# spatial features (csrc/spatial.cu) vs oracle/spatial_oracle.py
#   phases are float32 like the reference (np.angle of complex64): atan2f / cosf on the
#   device and in libm differ by a few ulp -> 5e-6 absolute on cos/sin/IPD (IPD compared
#   modulo 2 pi: a last-bit difference may wrap the other way), 5e-6 of max|.| for GCC/SRP;
#   MSC is float64 arithmetic on exact complex64 inputs -> 1e-9.
# ---------------------------------------------------------------------------
TOL_PHASE = 5e-6


def _wrap_err(a, b):
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(np.mod(d + np.pi, 2 * np.pi) - np.pi)))


def spatial_compare(device, S, obs, steer, pairs, topo, expect):
    """
    Run every spatial entry point on S (C x T x F complex64) / obs (C x F x T) and compare
    with `expect` (dict of arrays: the oracle's or the reference's).  Returns worst errors.
    """
    C, T, F = S.shape
    St = torch.from_numpy(S).to(device)
    res = {}
    res["ipd"] = _wrap_err(P.ipd(St[0], St[1], 0).cpu().numpy(), expect["ipd"])
    res["ipd_cos"] = float(np.max(np.abs(P.ipd(St[0], St[1], 1).cpu().numpy() - expect["ipd_cos"])))
    got = P.ipd(St[0], St[C - 1], 2).cpu().numpy()
    assert got.shape == (T, 2 * F) and got.dtype == np.float32
    res["ipd_cos_sin"] = float(np.max(np.abs(got - expect["ipd_cos_sin"])))
    ot = torch.from_numpy(obs).to(device)[None]
    res["df"] = float(np.max(np.abs(P.directional_feats(ot, torch.from_numpy(steer).to(device))[0]
                                    .cpu().numpy() - expect["df"])))
    res["df_given_pairs"] = float(np.max(np.abs(
        P.directional_feats(ot, torch.from_numpy(steer).to(device)[None], pairs=pairs)[0].cpu().numpy()
        - expect["df_given_pairs"])))
    for k in ("ipd", "ipd_cos", "ipd_cos_sin", "df", "df_given_pairs"):
        assert res[k] <= TOL_PHASE, (k, res[k])
    from oracle import spatial_oracle as sp

    def rel(got, ref):
        return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))

    gk = dict(num_bins=F, num_doa=37)
    om, tau = sp.tdoa_grid(0.07, **gk)
    res["gcc"] = rel(P.gcc_phat(St[0], St[1], om, tau).cpu().numpy(), expect["gcc"])
    om, tau = sp.tdoa_grid(-0.05, samp_doa=False, **gk)
    res["gcc_tdoa_raw"] = rel(P.gcc_phat(St[0], St[1], om, tau, normalize=False, apply_floor=False)
                              .cpu().numpy(), expect["gcc_tdoa_raw"])
    tau = np.cos(0.3 - np.linspace(0, np.pi * 2, 25)) * 0.1 / 343
    om = np.linspace(0, 16000 / 2, F) * 2 * np.pi
    res["gcc_diag"] = rel(P.gcc_phat(St[0], St[1], om, tau).cpu().numpy(), expect["gcc_diag"])
    if C == 2:
        om, tau = sp.tdoa_grid(topo[1] - topo[0], **gk)
        srp = P.gcc_phat(St[0], St[1], om, tau)
    else:
        srp = torch.zeros((T, 37), dtype=torch.float64, device=device)
        for i in range(C):
            for j in range(i + 1, C):
                om, tau = sp.tdoa_grid(topo[j] - topo[i], **gk)
                P.gcc_phat(St[i], St[j], om, tau, out=srp)
        srp = srp * 2 / (C * (C - 1))
    res["srp"] = rel(srp.cpu().numpy(), expect["srp"])
    for k in ("gcc", "gcc_tdoa_raw", "gcc_diag", "srp"):
        assert res[k] <= TOL_PHASE, (k, res[k])
    for ctx in (0, 1, 2):
        res[f"msc_ctx{ctx}"] = rel(P.msc(St, context=ctx).cpu().numpy(), expect[f"msc_ctx{ctx}"])
    res["msc_raw"] = rel(P.msc(St, context=1, normalize=False).cpu().numpy(), expect["msc_raw"])
    for k in ("msc_ctx0", "msc_ctx1", "msc_ctx2", "msc_raw"):
        assert res[k] <= 1e-9, (k, res[k])
    return res


def check_spatial(device, rng, C, N, frame_len=256, hop=128):
    """Every spatial kernel vs oracle/spatial_oracle.py on a seeded mixture."""
    from oracle import spatial_oracle as sp
    x = structured_audio(rng, 1, C, N)[0]
    obs = oracle_stft(x, frame_len, hop, True, "hann", dtype=np.complex64)           # C x F x T
    S = np.ascontiguousarray(np.transpose(obs, (0, 2, 1)))
    F = obs.shape[1]
    steer = np.exp(-1j * rng.uniform(0, 2 * np.pi, size=(C, F)))
    pairs = [(0, C - 1), (C - 1, 0)] if C > 2 else [(1, 0)]
    topo = [0.04 * i for i in range(C)]
    gk = dict(num_bins=F, num_doa=37)
    expect = {
        "ipd": sp.ipd(S[0], S[1]), "ipd_cos": sp.ipd(S[0], S[1], cos=True),
        "ipd_cos_sin": sp.ipd(S[0], S[C - 1], cos=True, sin=True),
        "df": sp.directional_feats(obs, steer),
        "df_given_pairs": sp.directional_feats(obs, steer, df_pair=pairs),
        "gcc": sp.gcc_phat_linear(S[0], S[1], 0.07, **gk),
        "gcc_tdoa_raw": sp.gcc_phat_linear(S[0], S[1], -0.05, normalize=False, apply_floor=False,
                                           samp_doa=False, **gk),
        "gcc_diag": sp.gcc_phat_diag(S[0], S[1], 0.3, 0.1, num_doas=25, num_bins=F),
        "srp": sp.srp_phat_linear(S, topo, **gk),
        "msc_raw": sp.msc(S, context=1, normalize=False),
    }
    for ctx in (0, 1, 2):
        expect[f"msc_ctx{ctx}"] = sp.msc(S, context=ctx)
    return spatial_compare(device, S, obs, steer, pairs, topo, expect)


def check_spatial_fixture(device, name):
    """The same entry points against the REFERENCE's own outputs (tests/golden/ref_spatial.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_spatial.npz"))
    fl, hop = (int(v) for v in g[name + "/cfg"])
    obs = oracle_stft(g[name + "/mix"], fl, hop, True, "hann", dtype=np.complex64)
    S = np.ascontiguousarray(np.transpose(obs, (0, 2, 1)))
    expect = {k.split("/", 1)[1]: g[k] for k in g.files if k.startswith(name + "/")}
    pairs = [tuple(int(v) for v in p) for p in g[name + "/df_pairs"]]
    return spatial_compare(device, S, obs, g[name + "/steer"], pairs, list(g[name + "/topo"]), expect)


def check_fixed_beamformers(device):
    """
    Geometry-based beamformers of setk_b200.libs.beamformer against the REFERENCE's own weights
    and enhanced STFTs (tests/golden/ref_fixed_bf.npz): weights are host constants (<= 1e-12),
    the enhanced STFT comes from setk_apply (complex64 storage: 2e-6).
    """
    import os
    from setk_b200.libs import beamformer as BF
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_fixed_bf.npz"))
    obs = oracle_stft(g["mix"], 256, 128, True, "hann", dtype=np.complex64)         # 4 x F x T
    F = obs.shape[1]
    topo = [0.0, 0.05, 0.1, 0.15]
    objs = {"lin_ds": BF.LinearDSBeamformer(topo), "lin_sd": BF.LinearSDBeamformer(topo),
            "cir_ds": BF.CircularDSBeamformer(0.05, 4), "cir_sd": BF.CircularSDBeamformer(0.05, 3, center=True)}
    worst = 0.0
    for name, obj in objs.items():
        doa = float(g[name + "/doa"])
        w = obj.weight(doa, F)
        assert w.shape == g[name + "/weight"].shape
        assert np.max(np.abs(w - g[name + "/weight"])) <= 1e-12 * max(1.0, np.max(np.abs(w))), name
        enh = obj.run(doa, torch.from_numpy(obs).to(device))
        assert torch.is_tensor(enh)
        err = bo.rel_inf(enh.cpu().numpy(), g[name + "/enh"])
        assert err <= 2e-6, (name, err)
        worst = max(worst, err)
    assert np.array_equal(objs["cir_sd"].distance_mat, g["cir_sd/distance_mat"])
    enh = BF.FixedBeamformer(g["lin_sd/weight"]).run(torch.from_numpy(obs).to(device)).cpu().numpy()
    assert bo.rel_inf(enh, g["fixed/enh"]) <= 2e-6
    assert np.allclose(BF.beam_pattern(g["lin_ds/weight"], g["pattern/sv"]), g["pattern/single"], rtol=0, atol=1e-12)
    multi = BF.beam_pattern(np.stack([g["lin_ds/weight"], g["lin_sd/weight"]]), g["pattern/sv"])
    assert isinstance(multi, list) and np.allclose(np.stack(multi), g["pattern/multi"], rtol=0, atol=1e-12)
    assert np.allclose(BF.diffuse_covar(F, np.abs(np.subtract.outer(topo, topo)), diag_eps=0.01), g["diffuse"],
                       rtol=0, atol=1e-15)
    import pytest
    with pytest.raises(ValueError):
        objs["lin_ds"].run(30.0, torch.from_numpy(obs[:3]).to(device))
    with pytest.raises(RuntimeError):
        BF.beam_pattern(g["lin_ds/weight"][:, :3], g["pattern/sv"])
    return worst


def check_wpd_fixture(device, name, enh_tol=1e-4, mask_mean_tol=1e-5):
    """
    libs.wpe.facted_wpd (setk_wpe_step -> setk_cgmm_stft -> setk_cov x2 -> setk_weights -> setk_apply)
    against the REFERENCE's facted_wpd output (tests/golden/ref_wpd.npz).  The enhanced spectrum is
    compared after per-bin phase alignment (the principal eigenvector's phase is LAPACK's in the
    reference, SURVEY.md finding 4).  Observed on the CPU execution model: 7e-7 on the spectrum,
    mask differences 1e-7 mean / 3e-5 max; the bounds leave room for the CGMM's ill-conditioned
    cells (an eigenvalue at its floor) without letting a wrong stage through.
    """
    import os
    from setk_b200.libs import utils
    from setk_b200.libs.wpe import facted_wpd
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_wpd.npz"))
    fl, hop, taps, delay, ctx, ci, wi = (int(v) for v in g[name + "/cfg"])
    obs = oracle_stft(g[name + "/mix"], fl, hop, True, "hann", dtype=np.complex64)      # C x F x T
    x = torch.from_numpy(np.ascontiguousarray(np.transpose(obs, (0, 2, 1)))).to(device)  # C x T x F
    tf_mask, enh = facted_wpd(x, cgmm_iters=ci, wpd_iters=wi, taps=taps, delay=delay, context=ctx)
    tf_mask, enh = tf_mask.cpu().numpy(), enh.cpu().numpy()
    ref_mask, ref_enh = g[name + "/tf_mask"], g[name + "/enh"]
    assert tf_mask.shape == ref_mask.shape and enh.shape == ref_enh.shape
    ea, _ = bo.align_phase(enh.T, ref_enh.T)                                              # per bin
    err = bo.rel_inf(ea, ref_enh.T)
    md = np.abs(tf_mask - ref_mask)
    assert err <= enh_tol, (name, err)
    assert float(md.mean()) <= mask_mean_tol and float((md > 1e-3).mean()) <= 0.01, (name, md.mean(), md.max())
    return err, float(md.mean()), float(md.max())
