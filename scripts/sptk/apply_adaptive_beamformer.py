#!/usr/bin/env python
# coding=utf-8
"""
Do mvdr/gevd/... adaptive beamformer

Drop-in for the reference's scripts/sptk/apply_adaptive_beamformer.py: same
positional arguments, flags and defaults (reference lines 183-259, StftParser
libs/opts.py:21-49), same outputs (<dst_dir>/<key>.wav, PCM-16), with the
per-utterance numerics running on libsetk_b200's CUDA kernels:

    fused STFT + covariances  ->  fp64 per-bin weight solve  ->  fused apply + iSTFT

The offline path (no --online.chunk-size, no --vad-proportion) does not process one
utterance per launch: setk_b200/batch_cli.py sorts a look-ahead window of the scp by
length, fills pinned PCM-16 staging buffers of up to --batch-size ragged utterances on
a reader thread, runs them through a two-lane HostBatchStreamer (copies overlap
kernels) and writes PCM-16 wav files on a writer thread.  --batch-size / --lookahead
are the only flags the reference does not have.  Under torchrun every rank takes the
keys rank::world of the scp (run.pl's nj shards).

Deviations from the reference, all bug fixes:
  * --itf-mask is read from args.itf_mask (the reference opens args.tgt_mask);
  * the online path works (the reference passes normalize= to run(..., ban=));
  * num_bins follows --round-power-of-two.
"""
import argparse
import math
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from setk_b200.engine import BEAMFORMERS, BeamformPipeline  # noqa: E402
from setk_b200.libs.beamformer import OnlineGevdBeamformer, OnlineMvdrBeamformer  # noqa: E402
from setk_b200.libs.data_handler import (NumpyReader, ScriptReader, WaveReader,  # noqa: E402
                                         WaveWriter)
from setk_b200.libs.opts import StftParser, strtobool  # noqa: E402
from setk_b200.libs.utils import (cmat_abs, default_device, get_logger, get_plan,  # noqa: E402
                                  inverse_stft, nextpow2)

logger = get_logger(__name__)
beamformers = BEAMFORMERS


def compute_vad_masks(spectrogram, proportion):
    """
    We ignore several minimum values and keep proportion*100% energy
    (reference lines 50-71).  spectrogram: F x T torch tensor.
    Return: vad_mask T x F (bool tensor), index
    """
    import torch
    energy_mat = cmat_abs(spectrogram)
    energy_vec = torch.sort(energy_mat.flatten())[0]
    filter_energy = torch.sum(energy_vec) * (1 - proportion)
    csum = torch.cumsum(energy_vec, 0)
    over = torch.nonzero(csum > filter_energy)
    if over.numel():
        index = int(over[0])
        threshold = energy_vec[index]
    else:
        index = energy_vec.numel()
        threshold = energy_vec[-1]
    return (energy_mat < threshold).transpose(0, 1), index


def do_online_beamform(beamformer, speech_mask, interf_mask, stft_mat, args):
    """
    Do online beamformer(gevd, mvdr)  (reference lines 25-47)
    speech_mask T x F, stft_mat N x F x T (torch).  Return F x T
    """
    import torch
    chunk_size = args.chunk_size
    beamformer.reset_stats(args.alpha)
    num_chunks = math.ceil(stft_mat.shape[-1] / chunk_size)
    enh_chunks = []
    for c in range(num_chunks):
        base = chunk_size * c
        mask_n = None if interf_mask is None else interf_mask[base:base + chunk_size]
        chunk = beamformer.run(speech_mask[base:base + chunk_size],
                               stft_mat[:, :, base:base + chunk_size].contiguous(),
                               mask_n=mask_n, ban=args.ban)
        enh_chunks.append(chunk)
    return torch.cat(enh_chunks, dim=-1)


def run(args):
    import torch
    stft_kwargs = {
        "frame_len": args.frame_len,
        "frame_hop": args.frame_hop,
        "window": args.window,
        "center": bool(args.center),  # false to comparable with kaldi
    }
    dev = default_device()
    online = args.chunk_size > 0
    batched = (not online) and not (0.5 < args.vad_proportion < 1) and args.batch_size > 0
    if "LOCAL_RANK" in os.environ and dev.type == "cuda":
        import torch as _t
        dev = _t.device("cuda", int(os.environ["LOCAL_RANK"]))
        _t.cuda.set_device(dev)
    wave_reader = WaveReader(args.wav_scp, sr=args.sr, raw_pcm16=batched)
    MaskReader = {"numpy": NumpyReader, "kaldi": ScriptReader}
    tgt_mask_reader = MaskReader[args.fmt](args.tgt_mask)
    itf_mask_reader = MaskReader[args.fmt](args.itf_mask) if args.itf_mask else None
    if itf_mask_reader is not None:
        logger.info(f"Using interfering masks from {args.itf_mask}")
    n_fft = nextpow2(args.frame_len) if args.round_power_of_two else args.frame_len
    num_bins = n_fft // 2 + 1
    if batched:
        from setk_b200.batch_cli import run_batched
        logger.info(f"Using offline {args.beamformer} beamformer, batches of <= {args.batch_size}")
        num_done = run_batched(args, wave_reader, tgt_mask_reader, itf_mask_reader, stft_kwargs,
                               num_bins, dev, logger)
        logger.info(f"Processed {num_done:d} utterances " + f"out of {len(wave_reader):d}")
        return
    if not online:
        logger.info(f"Using offline {args.beamformer} beamformer")
    else:
        if args.chunk_size < 32:
            raise RuntimeError(f"Seems chunk size({args.chunk_size:.2f}) " +
                               "too small for online beamformer")
        if args.beamformer not in ("mvdr", "gevd"):
            raise KeyError(f"online beamformer supports mvdr / gevd, got {args.beamformer}")
        online_bf = {"mvdr": OnlineMvdrBeamformer, "gevd": OnlineGevdBeamformer}[args.beamformer](
            num_bins, args.channels, args.alpha)
        logger.info(f"Using online {args.beamformer} beamformer, chunk size = {args.chunk_size:d}")

    pipes = {}

    def pipeline(num_channels, nsamps):
        p = pipes.get(num_channels)
        if p is None or p.plan.max_samples < nsamps:
            if p is not None:
                p.plan.close()
            p = BeamformPipeline(num_channels, beamformer=args.beamformer,
                                 round_power_of_two=bool(args.round_power_of_two),
                                 ban=bool(args.ban), pmwf_ref=args.pmwf_ref,
                                 rank1_appro=args.rank1_appro, post_masking=bool(args.mask),
                                 max_batch=1, max_samples=nsamps, device=dev, **stft_kwargs)
            pipes[num_channels] = p
        return p

    num_done = 0
    with WaveWriter(args.dst_dir, sr=args.sr) as writer:
        for key, samps in wave_reader:
            if key not in tgt_mask_reader:
                continue
            if samps.ndim == 1:
                samps = samps[None]
            power = np.linalg.norm(samps[0], 2)**2 / samps[0].size
            logger.info(f"Processing utterance {key}, " +
                        f"signal power {10 * np.log10(power + 1e-5):.2f}...")
            audio = torch.from_numpy(np.ascontiguousarray(samps)).to(dev)[None]   # 1 x C x N
            C, N = samps.shape
            # prefer T x F
            speech_mask = torch.from_numpy(np.asarray(tgt_mask_reader[key], dtype=np.float32)).to(dev)
            interf_mask = None
            if itf_mask_reader is not None:
                interf_mask = torch.from_numpy(
                    np.asarray(itf_mask_reader[key], dtype=np.float32)).to(dev)
            else:
                # constraint [0, 1]
                speech_mask = torch.clamp(speech_mask, max=1.0)
            # make sure speech_mask at shape T x F
            if speech_mask.shape[0] == num_bins and speech_mask.shape[1] != num_bins:
                speech_mask = speech_mask.transpose(0, 1).contiguous()
                if interf_mask is not None:
                    interf_mask = interf_mask.transpose(0, 1).contiguous()
            need_stft = online or (0.5 < args.vad_proportion < 1)
            stft_mat = None
            if need_stft:
                pl = get_plan(C, args.frame_len, args.frame_hop, bool(args.center),
                              bool(args.round_power_of_two), args.window, N, dev)
                stft_mat = pl.stft(audio)[0]                                       # N x F x T
            if 0.5 < args.vad_proportion < 1:
                vad_mask, n_filtered = compute_vad_masks(stft_mat[0], args.vad_proportion)
                logger.info(f"Filtering {n_filtered} TF-masks...")
                speech_mask = torch.where(vad_mask, torch.full_like(speech_mask, 1.0e-4),
                                          speech_mask)
                if interf_mask is not None:
                    interf_mask = torch.where(vad_mask, torch.full_like(interf_mask, 1.0e-4),
                                              interf_mask)
            try:
                if not online:
                    pipe = pipeline(C, N)
                    wave, status = pipe.run(audio, speech_mask[None],
                                            None if interf_mask is None else interf_mask[None],
                                            clip_mask=False, normalize=True)
                    BeamformPipeline.raise_for_status(status, [key])
                    samps_enh = wave[0].cpu().numpy()
                else:
                    stft_enh = do_online_beamform(online_bf, speech_mask, interf_mask, stft_mat,
                                                  args)
                    if args.mask:
                        stft_enh = stft_enh * speech_mask.transpose(0, 1)
                    norm = float(np.max(np.abs(samps)))
                    samps_enh = inverse_stft(stft_enh, norm=norm, transpose=False,
                                             **stft_kwargs).cpu().numpy()
            except np.linalg.LinAlgError:
                logger.error(f"Raise linalg error: {key}")
                continue
            writer.write(key, samps_enh)
            num_done += 1
    logger.info(f"Processed {num_done:d} utterances " + f"out of {len(wave_reader):d}")


def get_parser():
    parser = argparse.ArgumentParser(
        description="Command to run adaptive(mvdr/gevd/pmwf) beamformer",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Multi-channel wave scripts in kaldi format")
    parser.add_argument("tgt_mask", type=str,
                        help="Scripts of target masks in kaldi's archive or numpy's ndarray")
    parser.add_argument("dst_dir", type=str, help="Location to dump enhanced wave files")
    parser.add_argument("--itf-mask", type=str, default="",
                        help="Scripts of interfering masks in kaldi's archive or numpy's ndarray")
    parser.add_argument("--mask-format", dest="fmt", choices=["kaldi", "numpy"], default="kaldi",
                        help="Define format of masks, kaldi's archives or numpy's ndarray")
    parser.add_argument("--beamformer", type=str, default="mvdr", choices=beamformers,
                        help="Type of adaptive beamformer to apply")
    parser.add_argument("--pmwf-ref", type=int, default=-1,
                        help="Reference channel for PMWF beamformer")
    parser.add_argument("--sr", type=int, default=16000, help="Sample rate of the waveform")
    parser.add_argument("--ban", type=strtobool, default=False,
                        help="Do Blind Analytical Normalization (BAN) or not")
    parser.add_argument("--rank1-appro", type=str, default="", choices=["", "none", "eig", "gev"],
                        help="Weather to use rank1 approximation in PMWF")
    parser.add_argument("--post-masking", dest="mask", type=strtobool, default=False,
                        help="Masking enhanced spectrogram after beamforming or not")
    parser.add_argument("--vad-proportion", type=float, default=1,
                        help="Energy proportion to filter silence masks [0.5, 1]")
    parser.add_argument("--online.alpha", default=0.8, dest="alpha", type=float,
                        help="Remember coefficient when updating covariance matrix")
    parser.add_argument("--online.chunk-size", default=-1, type=int, dest="chunk_size",
                        help="If >= 64, using online beamformer instead")
    parser.add_argument("--online.channels", default=4, type=int, dest="channels",
                        help="Number of channels available")
    parser.add_argument("--batch-size", default=64, type=int,
                        help="[setk_b200] utterances per device batch of the offline path "
                        "(0: one utterance per launch, like the reference's loop)")
    parser.add_argument("--lookahead", default=256, type=int,
                        help="[setk_b200] utterances read ahead and sorted by length before batching")
    parser.add_argument("--reader-threads", default=8, type=int,
                        help="[setk_b200] loader threads of the batched path (file reads in scp order)")
    return parser


if __name__ == "__main__":
    run(get_parser().parse_args())
