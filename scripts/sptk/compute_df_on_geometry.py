#!/usr/bin/env python
# coding=utf-8
"""
Compute directional features (DF) for linear arrays, based on given steer vectors

Drop-in for the reference's scripts/sptk/compute_df_on_geometry.py (same positional
arguments, flags, defaults and output archive), with the STFT and the features on
libsetk_b200's CUDA kernels (setk_stft, setk_directional_feats).
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

from setk_b200.libs.data_handler import ArchiveWriter, ScpReader, SpectrogramReader  # noqa: E402
from setk_b200.libs.opts import StftParser  # noqa: E402
from setk_b200.libs.spatial import directional_feats  # noqa: E402
from setk_b200.libs.utils import get_logger  # noqa: E402

logger = get_logger(__name__)


def run(args):
    stft_kwargs = {
        "frame_len": args.frame_len,
        "frame_hop": args.frame_hop,
        "round_power_of_two": args.round_power_of_two,
        "window": args.window,
        "center": args.center,  # false to comparable with kaldi
        "transpose": False  # F x T
    }
    stft_reader = SpectrogramReader(args.wav_scp, **stft_kwargs)
    if args.utt2idx:
        utt2idx = ScpReader(args.utt2idx, value_processor=int)
        logger.info(f"Using --utt2idx={args.utt2idx}")
    else:
        utt2idx = None
        logger.info(f"Using --doa-idx={args.doa_idx}")

    df_pair = [tuple(map(int, p.split(","))) for p in args.df_pair.split(";")]
    if not len(df_pair):
        raise RuntimeError(f"Bad configurations with --pair {args.df_pair}")
    logger.info(f"Compute directional feature with {df_pair}")

    steer_vector = np.load(args.steer_vector)                  # A x M x F

    num_done = 0
    with ArchiveWriter(args.dup_ark, args.scp) as writer:
        for key in stft_reader.index_keys:
            if utt2idx is not None and key not in utt2idx:
                logger.warning(f"Missing utt2idx for utterance {key}")
                continue
            stft = stft_reader.stft(stft_reader.read(key), as_tensor=True)      # M x F x T on the device
            if utt2idx is None:
                idx = [int(v) for v in str(args.doa_idx).split(",")]
                dfs = [directional_feats(stft, steer_vector[i], df_pair=df_pair) for i in idx]
                if len(dfs) == 1:
                    df = dfs[0]
                else:
                    dfs = torch.stack(dfs)                                       # A' x T x F
                    df = dfs.transpose(0, 1).reshape(dfs.shape[1], -1)
            else:
                df = directional_feats(stft, steer_vector[utt2idx[key]], df_pair=df_pair)
            writer.write(key, df.cpu().numpy())
            num_done += 1
            if not num_done % 1000:
                logger.info(f"Processed {num_done:d} utterance...")
    logger.info(f"Processed {num_done:d} utterances over {len(stft_reader):d}")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Command to compute directional features for linear arrays, "
        "based on given steer vector. Also see scripts/sptk/compute_steer_vector.py",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Multi-Channel wave scripts in kaldi format")
    parser.add_argument("steer_vector", type=str,
                        help="Pre-computed steer vector in each directions (in shape A x M x F, A: number "
                        "of DoAs, M: microphone number, F: FFT bins)")
    parser.add_argument("dup_ark", type=str, help="Location to dump features (in ark format)")
    parser.add_argument("--utt2idx", type=str, default="",
                        help="utt2idx for index (between [0, A - 1]) of the DoA.")
    parser.add_argument("--doa-idx", type=str, default=0,
                        help="DoA index for all utterances if --utt2idx=\"\"")
    parser.add_argument("--scp", type=str, default="",
                        help="If assigned, generate corresponding feature scripts")
    parser.add_argument("--df-pair", type=str, default="0,1",
                        help="Microphone pairs for directional feature computation")
    args = parser.parse_args()
    run(args)
