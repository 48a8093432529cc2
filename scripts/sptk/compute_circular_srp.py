#!/usr/bin/env python
# coding=utf-8
"""
Compute SRP-PHAT angular spectrum for circular arrays (diagonal microphone pairs)

Drop-in for the reference's scripts/sptk/compute_circular_srp.py (same positional
arguments, flags, defaults and output archive), with the STFT and the GCC-PHAT of
every pair on libsetk_b200's CUDA kernels (setk_stft, setk_gcc_phat).
"""
import argparse
import os
import sys

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from setk_b200.libs.data_handler import ArchiveWriter, SpectrogramReader  # noqa: E402
from setk_b200.libs.opts import StftParser  # noqa: E402
from setk_b200.libs.spatial import gcc_phat_diag  # noqa: E402
from setk_b200.libs.utils import get_logger, nextpow2  # noqa: E402

logger = get_logger(__name__)


def run(args):
    srp_pair = [tuple(map(int, p.split(","))) for p in args.diag_pair.split(";")]
    if not len(srp_pair):
        raise RuntimeError(f"Bad configurations with --pair {args.diag_pair}")
    logger.info(f"Compute gcc with {srp_pair}")
    stft_kwargs = {
        "frame_len": args.frame_len,
        "frame_hop": args.frame_hop,
        "round_power_of_two": args.round_power_of_two,
        "window": args.window,
        "center": args.center,  # false to comparable with kaldi
        "transpose": True  # T x F
    }
    num_done = 0
    num_ffts = nextpow2(args.frame_len) if args.round_power_of_two else args.frame_len
    reader = SpectrogramReader(args.wav_scp, **stft_kwargs)
    with ArchiveWriter(args.srp_ark, args.scp) as writer:
        for key in reader.index_keys:
            stft_mat = reader.stft(reader.read(key), as_tensor=True)          # N x T x F on the device
            num_done += 1
            srp = []
            for (i, j) in srp_pair:
                srp.append(gcc_phat_diag(stft_mat[i], stft_mat[j], min(i, j) * np.pi * 2 / args.n, args.d,
                                         num_bins=num_ffts // 2 + 1, sr=args.sr, num_doas=args.num_doas))
            srp = torch.stack(srp).mean(dim=0)
            nan = int(torch.isnan(srp).sum())
            if nan:
                raise RuntimeError(f"Matrix {key} has nan ({nan:d}) items)")
            writer.write(key, srp.cpu().numpy())
            if not num_done % 1000:
                logger.info(f"Processed {num_done:d} utterances...")
    logger.info(f"Processd {len(reader):d} utterances done")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Command to compute SRP augular spectrum for circular arrays",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Rspecifier for multi-channel wave")
    parser.add_argument("srp_ark", type=str, help="Location to dump features")
    parser.add_argument("--scp", type=str, default="", help="If assigned, generate corresponding scripts")
    parser.add_argument("--n", type=int, default=6, help="Number of arrays")
    parser.add_argument("--d", type=float, default=0.07, help="Diameter of circular array")
    parser.add_argument("--diag-pair", type=str, default="0,3;1,4;2,5",
                        help="Compute gcc between those diagonal arrays")
    parser.add_argument("--sr", type=int, default=16000, help="Sample rate of input wave")
    parser.add_argument("--num-doas", type=int, default=121,
                        help="Number of DoA to sample between 0 and 2pi")
    args = parser.parse_args()
    run(args)
