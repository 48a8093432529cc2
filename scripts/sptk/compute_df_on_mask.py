#!/usr/bin/env python
# coding=utf-8
"""
Compute directional features using steer vector, based on TF-mask

Drop-in for the reference's scripts/sptk/compute_df_on_mask.py (same positional
arguments, flags, defaults and output archive).  Everything numeric runs on
libsetk_b200's CUDA kernels: STFT (setk_stft), mask-weighted covariance (setk_cov),
principal eigenvector as the steering vector (setk_weights, PEVD), directional
features (setk_directional_feats).
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

from setk_b200.libs.beamformer import compute_covar, solve_pevd  # noqa: E402
from setk_b200.libs.data_handler import (ArchiveWriter, NumpyReader, ScriptReader,  # noqa: E402
                                         SpectrogramReader)
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
    feat_reader = SpectrogramReader(args.wav_scp, **stft_kwargs)
    MaskReader = {"numpy": NumpyReader, "kaldi": ScriptReader}
    mask_reader = MaskReader[args.fmt](args.mask_scp)

    df_pair = [tuple(map(int, p.split(","))) for p in args.df_pair.split(";")]
    if not len(df_pair):
        raise RuntimeError(f"Bad configurations with --pair {args.df_pair}")
    logger.info(f"Compute directional feature with {df_pair}")

    num_done = 0
    with ArchiveWriter(args.dup_ark, args.scp) as writer:
        for key in feat_reader.index_keys:
            if key not in mask_reader:
                logger.warning(f"Missing TF-mask for utterance {key}")
                continue
            obs = feat_reader.stft(feat_reader.read(key), as_tensor=True)       # N x F x T on the device
            speech_masks = mask_reader[key]
            _, F, _ = obs.shape
            if speech_masks.shape[0] == F:                                       # make sure T x F
                speech_masks = np.transpose(speech_masks)
            speech_masks = np.minimum(speech_masks, 1)
            speech_covar = compute_covar(obs, torch.from_numpy(np.ascontiguousarray(speech_masks)))
            sv = solve_pevd(speech_covar)                                        # F x N
            df = directional_feats(obs, sv.transpose(0, 1), df_pair=df_pair)
            writer.write(key, df.cpu().numpy())
            num_done += 1
            if not num_done % 1000:
                logger.info(f"Processed {num_done:d} utterance...")
    logger.info(f"Processed {num_done:d} utterances over {len(feat_reader):d}")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Command to compute directional features for arbitrary arrays, "
        "based on estimated TF-masks",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Multi-Channel wave scripts in kaldi format")
    parser.add_argument("mask_scp", type=str,
                        help="Scripts of masks in kaldi's archive or numpy's ndarray")
    parser.add_argument("dup_ark", type=str, help="Location to dump features in kaldi's archives")
    parser.add_argument("--scp", type=str, default="",
                        help="If assigned, generate corresponding feature scripts")
    parser.add_argument("--mask-format", dest="fmt", choices=["kaldi", "numpy"], default="kaldi",
                        help="Define format of masks, in kaldi's archives or numpy's ndarray")
    parser.add_argument("--df-pair", type=str, default="0,1",
                        help="Microphone pairs for directional feature computation")
    args = parser.parse_args()
    run(args)
