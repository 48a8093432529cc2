"""
GPU tier: the drop-in CLI and the host-side paths around the kernels, on the real
libsetk_b200.so -- the batched feeder (ragged batches, PCM-16 in and out), the
energy-VAD mask path, the online (chunked) beamformers, the MPDR routes of
BeamformPipeline and the fused PCM-16 output, each against the oracle.
"""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import parity_cases as pc
from oracle import beamformer_oracle as bo
from oracle import stft_oracle as so

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "scripts", "sptk", "apply_adaptive_beamformer.py")
KW = dict(frame_len=512, frame_hop=256, center=True, window="hann", transpose=False)


def _write_wav(path, x):
    import scipy.io.wavfile as wavfile
    wavfile.write(path, 16000, so.pcm16_from_float(x.T))


def _corpus(tmp_path, n_utts, lengths, C=4, seed=0):
    """Ragged synthetic utterances as PCM-16 wav files + IRM masks (.npy); returns {key: (x, m)}."""
    from setk_b200 import synth
    data, wav_lines, mask_lines = {}, [], []
    for u in range(n_utts):
        N = int(lengths[u])
        x, m = synth.make_batch(1, C, N, device="cuda", first=seed + u)
        x = so.float_from_pcm16(so.pcm16_from_float(x[0].cpu().numpy()))   # what the wav holds
        m = m[0].cpu().numpy()
        key = f"utt{u:03d}"
        _write_wav(str(tmp_path / f"{key}.wav"), x)
        np.save(tmp_path / f"{key}.npy", m)
        wav_lines.append(f"{key} {tmp_path / (key + '.wav')}")
        mask_lines.append(f"{key} {tmp_path / (key + '.npy')}")
        data[key] = (x, m)
    (tmp_path / "wav.scp").write_text("\n".join(wav_lines) + "\n")
    (tmp_path / "mask.scp").write_text("\n".join(mask_lines) + "\n")
    return data


def _run_cli(tmp_path, dst, *flags):
    cmd = [sys.executable, CLI, "--frame-len", "512", "--frame-hop", "256", "--mask-format", "numpy",
           *flags, str(tmp_path / "wav.scp"), str(tmp_path / "mask.scp"), str(dst)]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stderr


def _read(dst, key):
    import scipy.io.wavfile as wavfile
    sr, out = wavfile.read(str(dst / f"{key}.wav"))
    assert sr == 16000 and out.dtype == np.int16
    return out


def test_cli_batched_ragged_vs_oracle(cuda, tmp_path):
    """72 ragged utterances through the batched feeder (3 batches of <= 32): PMWF is phase
    invariant, so every output file is comparable with the oracle sample by sample."""
    rng = np.random.default_rng(5)
    lengths = rng.integers(9000, 30000, size=72)
    lengths[3], lengths[40] = 30000, 9001
    data = _corpus(tmp_path, 72, lengths)
    dst = tmp_path / "pmwf"
    log = _run_cli(tmp_path, dst, "--beamformer", "pmwf-0", "--batch-size", "32", "--lookahead", "48")
    assert "Batched feeder: 72 utterances" in log
    worst = 0
    for key, (x, m) in data.items():
        out = _read(dst, key)
        T = so.num_frames(x.shape[1], 512, 256, True)
        assert out.shape == (256 * (T - 1),)
        y, _, _ = bo.enhance_utterance(x, m, kind="pmwf", beta=0)
        ref = so.pcm16_from_float(y)
        d = np.abs(out.astype(np.int64) - ref.astype(np.int64))
        assert d.max() <= 1 and np.mean(d > 0) <= 0.05, (key, d.max(), np.mean(d > 0))
        worst = max(worst, int(d.max()))
    # the same corpus one utterance per launch (the reference's loop shape): same files to 1 LSB
    dst1 = tmp_path / "pmwf1"
    _run_cli(tmp_path, dst1, "--beamformer", "pmwf-0", "--batch-size", "0")
    for key in list(data)[:12]:
        d = np.abs(_read(dst, key).astype(np.int64) - _read(dst1, key).astype(np.int64))
        assert d.max() <= 1


def test_cli_batched_mvdr_post_mask_and_itf(cuda, tmp_path):
    """MVDR + --post-masking + --itf-mask through the batched path equals the per-utterance path."""
    rng = np.random.default_rng(6)
    data = _corpus(tmp_path, 9, rng.integers(8000, 16000, size=9), seed=100)
    lines = []
    for key, (x, m) in data.items():
        np.save(tmp_path / f"{key}.itf.npy", (1 - np.minimum(m, 1)) * 0.9)
        lines.append(f"{key} {tmp_path / (key + '.itf.npy')}")
    (tmp_path / "itf.scp").write_text("\n".join(lines) + "\n")
    flags = ["--beamformer", "mvdr", "--post-masking", "true", "--itf-mask", str(tmp_path / "itf.scp")]
    _run_cli(tmp_path, tmp_path / "b", *flags, "--batch-size", "4")
    _run_cli(tmp_path, tmp_path / "s", *flags, "--batch-size", "0")
    for key in data:
        a, b = _read(tmp_path / "b", key), _read(tmp_path / "s", key)
        assert a.shape == b.shape and np.abs(a).max() > 500
        assert np.abs(a.astype(np.int64) - b.astype(np.int64)).max() <= 1


def test_cli_vad_proportion_vs_oracle(cuda, tmp_path):
    """--vad-proportion 0.9 (apply_adaptive_beamformer.py:50-71): sort of all |X_0|, cumulative
    energy threshold, masked cells set to 1e-4."""
    data = _corpus(tmp_path, 2, [20000, 14000], seed=200)
    dst = tmp_path / "vad"
    log = _run_cli(tmp_path, dst, "--beamformer", "pmwf-0", "--vad-proportion", "0.9")
    assert "Filtering" in log
    for key, (x, m) in data.items():
        y, _, _ = bo.enhance_utterance(x, m, kind="pmwf", beta=0, vad_proportion=0.9)
        ref = so.pcm16_from_float(y)
        d = np.abs(_read(dst, key).astype(np.int64) - ref.astype(np.int64))
        assert d.max() <= 2 and np.mean(d > 0) <= 0.05, (key, d.max(), np.mean(d > 0))


def _load_cli_module():
    spec = importlib.util.spec_from_file_location("cli_adaptive", CLI)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("kind", ["mvdr", "gevd"])
def test_online_beamformer_vs_oracle(cuda, kind, tmp_path):
    """--online.chunk-size 64 (apply_adaptive_beamformer.py:25-47 with beamformer.py:286-320,
    685-728): every chunk against the oracle's recursion after per-bin phase alignment, and the
    CLI run end to end."""
    from setk_b200 import synth
    from setk_b200.libs import beamformer as BF
    from setk_b200.libs.utils import set_default_device
    set_default_device("cuda")
    cli = _load_cli_module()
    x, m = synth.make_batch(1, 4, 48000, device=cuda, first=300)
    xs = x[0].cpu().numpy()
    S = so.multichannel_stft(xs, round_power_of_two=True, **KW)             # C x F x T complex128
    mask = np.minimum(m[0].cpu().numpy().astype(np.float64), 1)
    T = S.shape[-1]
    args = cli.get_parser().parse_args(["a", "b", "c", "--online.chunk-size", "64",
                                        "--online.alpha", "0.8", "--beamformer", kind])
    cls = {"mvdr": BF.OnlineMvdrBeamformer, "gevd": BF.OnlineGevdBeamformer}[kind]
    bf = cls(257, 4, 0.8)
    stft_dev = torch.from_numpy(S.astype(np.complex64)).to(cuda)
    enh = cli.do_online_beamform(bf, torch.from_numpy(mask.astype(np.float32)).to(cuda), None,
                                 stft_dev, args).cpu().numpy()
    assert enh.shape == (257, T)
    st = bo.OnlineState(257, 4, 0.8)
    for c0 in range(0, T, 64):
        ref = st.run(kind, mask[c0:c0 + 64], S[:, :, c0:c0 + 64].astype(np.complex64).astype(np.complex128))
        got = enh[:, c0:c0 + 64]
        # per-bin phase: one unit scalar per bin and chunk (the eigenvector's, SURVEY finding 4)
        al, _ = bo.align_phase(got, ref)
        assert bo.rel_inf(al, ref) <= 2e-4, (kind, c0, bo.rel_inf(al, ref))
    # and the CLI path itself runs end to end on the device
    _write_wav(str(tmp_path / "u.wav"), xs)
    np.save(tmp_path / "u.npy", m[0].cpu().numpy())
    (tmp_path / "wav.scp").write_text(f"u {tmp_path / 'u.wav'}\n")
    (tmp_path / "mask.scp").write_text(f"u {tmp_path / 'u.npy'}\n")
    _run_cli(tmp_path, tmp_path / "on", "--beamformer", kind, "--online.chunk-size", "64")
    out = _read(tmp_path / "on", "u")
    assert out.shape == (256 * (T - 1),) and np.abs(out).max() > 500


@pytest.mark.parametrize("kind", ["mpdr", "mpdr-whiten"])
def test_pipeline_mpdr_routes(cuda, kind):
    """BeamformPipeline("mpdr" / "mpdr-whiten"): Ry from the all-ones mask (beamformer.py:555-590)."""
    from setk_b200 import synth
    x, m = synth.make_batch(2, 4, 32000, device=cuda, first=400)
    err = pc.mvdr_end_to_end(cuda, x.cpu().numpy(), m.cpu().numpy(), kind=kind)
    assert err <= pc.TOL_E2E, (kind, err)


def test_fused_pcm16_output_is_bit_identical(cuda):
    """setk_apply_istft_pcm16 == setk_apply_istft + setk_float_to_pcm16, ragged batch included."""
    from setk_b200 import synth
    from setk_b200 import plan as P
    from setk_b200.engine import BeamformPipeline
    x, m = synth.make_batch(3, 4, 24000, device=cuda, first=500)
    ns = torch.tensor([24000, 17000, 9000], dtype=torch.int32, device=cuda)
    pipe = BeamformPipeline(4, "mvdr", max_batch=3, max_samples=24000, device=cuda)
    for n_samples in (None, ns):
        wave, st = pipe.run(x, m, n_samples=n_samples)
        pcm, st2 = pipe.run(x, m, n_samples=n_samples, pcm16_out=True)
        assert int(st.abs().sum()) == 0 and pcm.dtype == torch.int16
        assert torch.equal(P.float_to_pcm16(wave), pcm)
    wave, _ = pipe.run(x, m, normalize=False)
    pcm, _ = pipe.run(x, m, normalize=False, pcm16_out=True)
    assert torch.equal(P.float_to_pcm16(wave), pcm)


def test_pmwf_two_streams_do_not_share_scratch(cuda):
    """PMWF with automatic reference selection on two streams at once (the scratch
    used to be one process-global buffer)."""
    from setk_b200 import _lib, plan as P
    rng = np.random.default_rng(9)
    Rs = [torch.from_numpy(np.stack([pc.random_rank1_plus(rng, 257, 4) for _ in range(8)])).to(cuda)
          for _ in range(2)]
    Rn = [torch.from_numpy(np.stack([pc.random_hpd(rng, 257, 4) for _ in range(8)])).to(cuda)
          for _ in range(2)]
    ref = [P.weights(_lib.BF_PMWF, Rs[i], Rn[i], beta=float(i), ref_channel=-1) for i in range(2)]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda) for _ in range(2)]
    for rep in range(10):
        outs = []
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                outs.append(P.weights(_lib.BF_PMWF, Rs[i], Rn[i], beta=float(i), ref_channel=-1))
        torch.cuda.synchronize()
        for i in range(2):
            assert torch.equal(outs[i][0], ref[i][0]) and torch.equal(outs[i][2], ref[i][2])



def test_streamer_compressed_masks(cuda):
    """HostBatchStreamer(cm_masks=True): PCM-16 samples + Kaldi CompressedMatrix masks from pinned host
    buffers.  The device expansion is bit-identical to the archive reader, so the output equals the
    same streamer fed the expanded float32 masks bit for bit, and the oracle run on the expanded masks
    (what the reference computes from such an archive) within the end-to-end tolerance."""
    import io
    from setk_b200 import synth
    from setk_b200 import plan as P
    from setk_b200.engine import BeamformPipeline, HostBatchStreamer
    from setk_b200.libs.data_handler import read_kaldi_matrix
    B, C, N = 3, 4, 24000
    x, m = synth.make_batch(B, C, N, device=cuda, first=700)
    pcm = P.float_to_pcm16(x)
    T, F = m.shape[1], m.shape[2]
    blobs = pc.pack_cm_blobs(list(m.cpu().numpy()), T, F)
    assert blobs.shape[1] == HostBatchStreamer.cm_slot_bytes(T, F)
    m_dec = np.stack([read_kaldi_matrix(io.BytesIO(b"\0BCM " + blobs[b, :16 + F * (8 + T)].tobytes()))
                      for b in range(B)])
    h_pcm = pcm.cpu().pin_memory()
    h_cm = torch.from_numpy(blobs).pin_memory()
    h_m = torch.from_numpy(m_dec).pin_memory()

    def mk():
        return BeamformPipeline(C, "mvdr", max_batch=B, max_samples=N, device=cuda)

    n_out = mk().run(x, m)[0].shape[1]
    outs = []
    for cm, hm in ((True, h_cm), (False, h_m)):
        st = HostBatchStreamer(mk, B, C, N, slots=2, pcm16=True, pcm16_out=True, device=cuda, cm_masks=cm)
        h_out = torch.empty((B, n_out), dtype=torch.int16).pin_memory()
        lane = st.submit(h_pcm, hm, h_out)
        st.synchronize()
        assert int(lane["status"].abs().sum()) == 0
        if cm:
            assert int(lane["cm_status"].abs().sum()) == 0
        outs.append(h_out.clone())
    assert torch.equal(outs[0], outs[1])
    # against the oracle on the expanded masks (the reference's view of the archive)
    xf = (pcm.cpu().numpy().astype(np.float32) / 32768.0)
    err = pc.mvdr_end_to_end(cuda, xf, m_dec, kind="mvdr")
    assert err <= pc.TOL_E2E, err
