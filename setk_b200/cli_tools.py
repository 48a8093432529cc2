"""
setk_b200.cli_tools -- the feature / WPD command lines behind scripts/sptk/
{compute_ipd_and_linear_srp, compute_df_on_geometry, compute_circular_srp, apply_wpd}.py.

Each keeps the positional arguments, flag names, destinations and defaults of the reference
script of the same name (scripts/sptk/*.py of funcwj/setk; cited per function) so that recipes
written against the reference keep working; everything numeric runs on libsetk_b200's CUDA
kernels through setk_b200.libs.  Flags are declared as tables (name, type, default, dest, help).
"""
import argparse

import numpy as np
import torch

from .libs.data_handler import ArchiveWriter, ScpReader, SpectrogramReader, WaveWriter
from .libs.opts import StftParser, str2tuple, strtobool
from .libs.utils import get_logger, inverse_stft, nextpow2

logger = get_logger(__name__)


def _parser(description, positionals, flags):
    p = argparse.ArgumentParser(description=description,
                                formatter_class=argparse.ArgumentDefaultsHelpFormatter,
                                parents=[StftParser.parser])
    for name, text in positionals:
        p.add_argument(name, type=str, help=text)
    for name, typ, default, dest, text in flags:
        kw = dict(type=typ, default=default, help=text)
        if dest:
            kw["dest"] = dest
        p.add_argument(name, **kw)
    return p


def _stft_options(args, transpose):
    return dict(frame_len=args.frame_len, frame_hop=args.frame_hop, window=args.window,
                center=args.center, round_power_of_two=args.round_power_of_two, transpose=transpose)


def _bins(args):
    n = nextpow2(args.frame_len) if args.round_power_of_two else args.frame_len
    return n // 2 + 1


def _pairs(text, flag):
    pairs = []
    for item in text.split(";"):
        idx = [int(v) for v in item.split(",")]
        if len(idx) != 2:
            raise ValueError(f"Invalid {flag} configuration detected: {text}")
        pairs.append(tuple(idx))
    if not pairs:
        raise RuntimeError(f"Bad configurations with {flag} {text}")
    return pairs


def _each_stft(reader):
    """(key, samples, STFT as a device tensor) for every utterance of a SpectrogramReader."""
    for key in reader.index_keys:
        samps = reader.read(key)
        yield key, samps, reader.stft(samps, as_tensor=True)


def _dump_archive(args_ark, args_scp, items, what):
    done = 0
    with ArchiveWriter(args_ark, args_scp) as writer:
        for key, feats in items:
            writer.write(key, feats.cpu().numpy())
            done += 1
            if done % 1000 == 0:
                logger.info(f"Processed {done} utterances...")
    logger.info(f"Processed {what} for {done} utterances")


# ---------------------------------------------------------------------------------------------
# compute_ipd_and_linear_srp.py (reference: compute_spatial_feats :21-50, run :53-73, flags :76-143)
# ---------------------------------------------------------------------------------------------
def spatial_feats_main(argv=None):
    from .libs.spatial import ipd, msc, srp_phat_linear
    args = _parser(
        "Spatial features of a multi-channel recording: SRP-PHAT angular spectrum of a linear array "
        "(srp), magnitude squared coherence (msc) or inter-channel phase differences (ipd)",
        [("wav_scp", "multi-channel wave script (Kaldi format)"),
         ("dup_ark", "Kaldi archive the features are written to")],
        [("--scp", str, "", None, "also write the archive's script file here"),
         ("--type", str, "srp", None, "srp | msc | ipd"),
         ("--srp.sample-rate", int, 16000, "samp_frequency", "sample rate of the input"),
         ("--srp.sample-tdoa", strtobool, False, "samp_tdoa", "sample the TDoA axis instead of the DoA axis"),
         ("--srp.num_doa", int, 181, "num_doa", "directions sampled between 0 and 180 degrees"),
         ("--srp.topo", str2tuple, "0,0.2,0.4,0.8", "linear_topo", "microphone positions of the linear array"),
         ("--ipd.cos", strtobool, False, "ipd_cos", "cosIPD instead of IPD"),
         ("--ipd.sin", strtobool, False, "ipd_sin", "append sinIPD to cosIPD"),
         ("--ipd.pair", str, "0,1", "ipd_pair", "channel pairs, e.g. 0,3;1,4"),
         ("--msc.ctx", int, 1, "msc_ctx", "frame context of the coherence estimate")]).parse_args(argv)
    if args.type not in ("srp", "msc", "ipd"):
        raise SystemExit(f"--type: invalid choice {args.type!r}")

    def features(S):                                   # S: N x T x F on the device
        if args.type == "srp":
            return srp_phat_linear(S, args.linear_topo, sample_frequency=args.samp_frequency,
                                   num_doa=args.num_doa, num_bins=_bins(args), samp_doa=not args.samp_tdoa)
        if args.type == "msc":
            return msc(S, context=args.msc_ctx)
        if S.ndim < 3:
            raise ValueError("Only one-channel STFT available")
        cols = []
        for left, right in _pairs(args.ipd_pair, "--ipd.pair"):
            if right > S.shape[0]:
                raise RuntimeError(f"Could not access channel {right}")
            cols.append(ipd(S[left], S[right], cos=args.ipd_cos, sin=args.ipd_sin))
        return torch.cat(cols, dim=1)

    reader = SpectrogramReader(args.wav_scp, **_stft_options(args, transpose=True))
    _dump_archive(args.dup_ark, args.scp, ((k, features(S)) for k, _, S in _each_stft(reader)),
                  args.type.upper())


# ---------------------------------------------------------------------------------------------
# compute_df_on_geometry.py (reference: run :18-71, flags :74-112)
# ---------------------------------------------------------------------------------------------
def df_on_geometry_main(argv=None):
    from .libs.spatial import directional_feats
    args = _parser(
        "Directional features of a linear array for given steering vectors",
        [("wav_scp", "multi-channel wave script (Kaldi format)"),
         ("steer_vector", "steering vectors, .npy of shape A x M x F (directions x microphones x bins)"),
         ("dup_ark", "Kaldi archive the features are written to")],
        [("--utt2idx", str, "", None, "per-utterance direction index in [0, A - 1]"),
         ("--doa-idx", str, 0, None, "direction index (or comma list) for all utterances without --utt2idx"),
         ("--scp", str, "", None, "also write the archive's script file here"),
         ("--df-pair", str, "0,1", None, "microphone pairs, e.g. 0,1;0,2")]).parse_args(argv)
    pairs = _pairs(args.df_pair, "--df-pair")
    table = np.load(args.steer_vector)
    utt2idx = ScpReader(args.utt2idx, value_processor=int) if args.utt2idx else None
    logger.info(f"Compute directional feature with {pairs}")

    def items():
        reader = SpectrogramReader(args.wav_scp, **_stft_options(args, transpose=False))
        for key, _, S in _each_stft(reader):           # S: M x F x T
            if utt2idx is not None:
                if key not in utt2idx:
                    logger.warning(f"Missing utt2idx for utterance {key}")
                    continue
                yield key, directional_feats(S, table[utt2idx[key]], df_pair=pairs)
                continue
            chosen = [int(v) for v in str(args.doa_idx).split(",")]
            per_dir = [directional_feats(S, table[i], df_pair=pairs) for i in chosen]
            if len(per_dir) == 1:
                yield key, per_dir[0]
            else:                                       # T x (directions * F)
                stacked = torch.stack(per_dir)
                yield key, stacked.transpose(0, 1).reshape(stacked.shape[1], -1)

    _dump_archive(args.dup_ark, args.scp, items(), "DF")


# ---------------------------------------------------------------------------------------------
# compute_circular_srp.py (reference: run :17-58, flags :61-92)
# ---------------------------------------------------------------------------------------------
def circular_srp_main(argv=None):
    from .libs.spatial import gcc_phat_diag
    args = _parser(
        "SRP angular spectrum of a circular array from its diagonal microphone pairs",
        [("wav_scp", "multi-channel wave script (Kaldi format)"),
         ("srp_ark", "Kaldi archive the spectra are written to")],
        [("--scp", str, "", None, "also write the archive's script file here"),
         ("--n", int, 6, None, "microphones on the circle"),
         ("--d", float, 0.07, None, "diameter of the array in metres"),
         ("--diag-pair", str, "0,3;1,4;2,5", None, "diagonal pairs whose GCC-PHAT is averaged"),
         ("--sr", int, 16000, None, "sample rate of the input"),
         ("--num-doas", int, 121, None, "directions sampled between 0 and 2 pi")]).parse_args(argv)
    pairs = _pairs(args.diag_pair, "--diag-pair")
    logger.info(f"Compute gcc with {pairs}")

    def items():
        reader = SpectrogramReader(args.wav_scp, **_stft_options(args, transpose=True))
        for key, _, S in _each_stft(reader):           # S: N x T x F
            acc = None
            for i, j in pairs:
                g = gcc_phat_diag(S[i], S[j], min(i, j) * np.pi * 2 / args.n, args.d, num_bins=_bins(args),
                                  sr=args.sr, num_doas=args.num_doas)
                acc = g if acc is None else acc + g
            srp = acc / len(pairs)
            bad = int(torch.isnan(srp).sum())
            if bad:
                raise RuntimeError(f"Matrix {key} has nan ({bad:d}) items)")
            yield key, srp

    _dump_archive(args.srp_ark, args.scp, items(), "SRP")


# ---------------------------------------------------------------------------------------------
# apply_wpd.py (reference: run :20-59, flags :62-109)
# ---------------------------------------------------------------------------------------------
def wpd_main(argv=None):
    from .libs.wpe import facted_wpd
    args = _parser(
        "Joint dereverberation and denoising with the factored WPD beamformer",
        [("wav_scp", "multi-channel wave script (Kaldi format)"),
         ("dst_dir", "directory the enhanced audio is written to")],
        [("--taps", int, 10, None, "prediction taps of the WPE stage"),
         ("--delay", int, 3, None, "prediction delay of the WPE stage"),
         ("--context", int, 1, None, "frame context of the first variance estimate"),
         ("--wpd-iters", int, 3, None, "WPD iterations"),
         ("--cgmm-iters", int, 20, None, "CGMM iterations inside every WPD iteration"),
         ("--update-alpha", strtobool, False, None, "update the CGMM priors in the M-step"),
         ("--sr", int, 16000, None, "sample rate of the input"),
         ("--dump-mask", strtobool, False, None, "also save the speech mask as <dst_dir>/<key>.npy")]
    ).parse_args(argv)
    opts = _stft_options(args, transpose=True)
    reader = SpectrogramReader(args.wav_scp, **opts)
    opts.pop("round_power_of_two")
    done = 0
    with WaveWriter(args.dst_dir, sr=args.sr) as writer:
        for key, samps, obs in _each_stft(reader):      # obs: N x T x F
            logger.info(f"Processing utt {key}...")
            if obs.ndim != 3:
                raise RuntimeError(f"Expected 3D array, but got {obs.ndim}")
            try:
                tf_mask, enh = facted_wpd(obs, wpd_iters=args.wpd_iters, cgmm_iters=args.cgmm_iters,
                                          update_alpha=args.update_alpha, context=args.context,
                                          taps=args.taps, delay=args.delay)
            except np.linalg.LinAlgError:
                logger.warning(f"{key}: Failed cause LinAlgError in wpd")
                continue
            wave = inverse_stft(enh, norm=float(np.max(np.abs(samps))), **opts)
            writer.write(key, wave.cpu().numpy())
            if args.dump_mask:
                np.save(f"{args.dst_dir}/{key}", tf_mask[..., 0].cpu().numpy())
            done += 1
            if done % 100 == 0:
                logger.info(f"Processed {done:d} utterances...")
    logger.info(f"Processed {done:d} utterances over {len(reader):d}")


# ---------------------------------------------------------------------------------------------
# compute_df_on_mask.py (reference: run :21-59, flags :62-95)
# ---------------------------------------------------------------------------------------------
def df_on_mask_main(argv=None):
    from .libs.beamformer import compute_covar, solve_pevd
    from .libs.data_handler import NumpyReader, ScriptReader
    from .libs.spatial import directional_feats
    args = _parser(
        "Directional features of an arbitrary array; the steering vector is the principal eigenvector "
        "of the mask-weighted spatial covariance",
        [("wav_scp", "multi-channel wave script (Kaldi format)"),
         ("mask_scp", "script of the TF-masks (Kaldi archive entries or .npy files)"),
         ("dup_ark", "Kaldi archive the features are written to")],
        [("--scp", str, "", None, "also write the archive's script file here"),
         ("--mask-format", str, "kaldi", "fmt", "kaldi | numpy"),
         ("--df-pair", str, "0,1", None, "microphone pairs, e.g. 0,1;0,2")]).parse_args(argv)
    if args.fmt not in ("kaldi", "numpy"):
        raise SystemExit(f"--mask-format: invalid choice {args.fmt!r}")
    pairs = _pairs(args.df_pair, "--df-pair")
    masks = (NumpyReader if args.fmt == "numpy" else ScriptReader)(args.mask_scp)
    logger.info(f"Compute directional feature with {pairs}")

    def items():
        reader = SpectrogramReader(args.wav_scp, **_stft_options(args, transpose=False))
        for key, _, obs in _each_stft(reader):          # obs: N x F x T
            if key not in masks:
                logger.warning(f"Missing TF-mask for utterance {key}")
                continue
            m = masks[key]
            if m.shape[0] == obs.shape[1]:              # given as F x T
                m = np.transpose(m)
            m = torch.from_numpy(np.ascontiguousarray(np.minimum(m, 1)))
            steer = solve_pevd(compute_covar(obs, m))   # F x N
            yield key, directional_feats(obs, steer.transpose(0, 1), df_pair=pairs)

    _dump_archive(args.dup_ark, args.scp, items(), "DF")
