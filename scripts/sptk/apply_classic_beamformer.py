#!/usr/bin/env python
# coding=utf-8
"""
Apply a classic beamformer (delay-and-sum / super-directive, linear & circular array)

Drop-in for the reference's scripts/sptk/apply_classic_beamformer.py (same positional
arguments, flags, defaults and outputs: <dst_dir>/<key>.wav, PCM-16).  The weights come
from the array geometry (host constants, libs/beamformer.py:380-512); the STFT, the
beamforming and the inverse STFT run on libsetk_b200's CUDA kernels (setk_stft,
setk_apply, setk_istft).
"""
import argparse
import math
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from setk_b200.libs.beamformer import (CircularDSBeamformer, CircularSDBeamformer,  # noqa: E402
                                       LinearDSBeamformer, LinearSDBeamformer)
from setk_b200.libs.data_handler import ScpReader, SpectrogramReader, WaveWriter  # noqa: E402
from setk_b200.libs.opts import StftParser, str2tuple, strtobool  # noqa: E402
from setk_b200.libs.utils import check_doa, get_logger, inverse_stft  # noqa: E402

logger = get_logger(__name__)
beamformers = ["ds", "sd"]


def do_online_beamform(beamformer, doa, stft_mat, args):
    """One DoA per chunk of --chunk-len frames (apply_classic_beamformer.py:20-30)."""
    chunk_size = args.chunk_len
    enh_chunks = []
    for c in range(len(doa)):
        base = chunk_size * c
        enh_chunks.append(beamformer.run(doa[c], stft_mat[:, :, base:base + chunk_size].contiguous(),
                                         c=args.speed, sr=args.sr))
    return torch.cat(enh_chunks, dim=1)


def process_doa(doa, online):
    if online:
        return list(map(float, doa if isinstance(doa, (list, tuple)) else [doa]))
    # --utt2doa lines arrive as token lists (num_tokens=-1); the reference's float(list)
    # raises there (apply_classic_beamformer.py:36) -- take the single value instead
    return float(doa[0] if isinstance(doa, (list, tuple)) else doa)


def parse_doa(args, online):
    if args.utt2doa:
        reader = ScpReader(args.utt2doa, value_processor=lambda doa: process_doa(doa, online),
                           num_tokens=-1)
        utt2doa = reader.get
        logger.info(f"Use --utt2doa={args.utt2doa} for each utterance")
    else:
        doa = process_doa(args.doa, online)
        utt2doa = lambda _: doa  # noqa: E731
        logger.info(f"Use --doa={doa} for all utterances")
    return utt2doa


def run(args):
    stft_kwargs = {
        "frame_len": args.frame_len,
        "frame_hop": args.frame_hop,
        "window": args.window,
        "center": args.center,
        "transpose": False
    }
    if args.geometry == "linear":
        cls = LinearDSBeamformer if args.beamformer == "ds" else LinearSDBeamformer
        beamformer = cls(linear_topo=args.linear_topo)
    else:
        cls = CircularDSBeamformer if args.beamformer == "ds" else CircularSDBeamformer
        beamformer = cls(radius=args.circular_radius, num_arounded=args.circular_around,
                         center=args.circular_center)
    online = args.chunk_len > 0
    utt2doa = parse_doa(args, online)
    spectrogram_reader = SpectrogramReader(args.wav_scp, round_power_of_two=args.round_power_of_two,
                                           **stft_kwargs)
    done = 0
    with WaveWriter(args.dst_dir, sr=args.sr) as writer:
        for key in spectrogram_reader.index_keys:
            doa = utt2doa(key)
            if doa is None:
                logger.info(f"Missing doa for utterance {key}")
                continue
            if not check_doa(args.geometry, doa, online):
                logger.info(f"Invalid doa {doa} for utterance {key}")
                continue
            samps_in = spectrogram_reader.read(key)
            stft_src = spectrogram_reader.stft(samps_in, as_tensor=True)           # N x F x T on the device
            if online:
                num_chunks = math.ceil(stft_src.shape[-1] / args.chunk_len)
                if len(doa) != num_chunks:
                    mn = math.ceil(stft_src.shape[-1] / len(doa))
                    mx = math.floor(stft_src.shape[-1] / (len(doa) - 1)) if len(doa) > 1 else mn
                    logger.info(f"Invalid chunk length {args.chunk_len} for utterance {key},"
                                f" expected --chunk-len from {mn} to {mx}")
                    continue
                stft_enh = do_online_beamform(beamformer, doa, stft_src, args)
            else:
                stft_enh = beamformer.run(doa, stft_src, c=args.speed, sr=args.sr)
            norm = float(np.max(np.abs(samps_in))) if args.normalize else None
            samps = inverse_stft(stft_enh, **stft_kwargs, norm=norm)
            writer.write(key, samps.cpu().numpy())
            done += 1
    logger.info(f"Processed {done} utterances over {len(spectrogram_reader)}")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Command to apply classic beamformer (linear & circular array).",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Rspecifier for multi-channel wave file")
    parser.add_argument("dst_dir", type=str, help="Directory to dump enhanced results")
    parser.add_argument("--beamformer", type=str, default="ds", choices=beamformers,
                        help="Type of classic beamformer to apply")
    parser.add_argument("--sr", type=int, default=16000, help="Sample rate of the input wave")
    parser.add_argument("--speed", type=float, default=343, help="Speed of sound")
    parser.add_argument("--geometry", type=str, choices=["linear", "circular"], default="linear",
                        help="Geometry of the microphone array")
    parser.add_argument("--linear-topo", type=str2tuple, default=(),
                        help="Topology of linear microphone arrays")
    parser.add_argument("--circular-around", type=int, default=6,
                        help="Number of the micriphones in circular arrays")
    parser.add_argument("--circular-radius", type=float, default=0.05, help="Radius of circular array")
    parser.add_argument("--circular-center", type=strtobool, default=False,
                        help="Is there a microphone put in the center of the circular array?")
    parser.add_argument("--utt2doa", type=str, default="",
                        help="Given DoA for each utterances, in degrees")
    parser.add_argument("--doa", type=str, default="0",
                        help="DoA for all utterances if --utt2doa is not assigned")
    parser.add_argument("--normalize", type=strtobool, default=False,
                        help="Normalize stft after enhancement?")
    parser.add_argument("--chunk-len", type=int, default=-1,
                        help="Number frames per chunk (for online setups)")
    args = parser.parse_args()
    run(args)
