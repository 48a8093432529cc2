#!/usr/bin/env python
# coding=utf-8
"""
Joint dereverberation & denoising (factored form of WPD)

Drop-in for the reference's scripts/sptk/apply_wpd.py (same positional arguments, flags,
defaults and outputs: <dst_dir>/<key>.wav PCM-16, optionally <dst_dir>/<key>.npy with the
speech mask), with the STFT, every stage of facted_wpd (libs/wpe.py:113-177: WPE step, CGMM,
covariances, MVDR solve, beamforming) and the inverse STFT on libsetk_b200's CUDA kernels.
"""
import argparse
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from setk_b200.libs.data_handler import SpectrogramReader, WaveWriter  # noqa: E402
from setk_b200.libs.opts import StftParser, strtobool  # noqa: E402
from setk_b200.libs.utils import get_logger, inverse_stft  # noqa: E402
from setk_b200.libs.wpe import facted_wpd  # noqa: E402

logger = get_logger(__name__)


def run(args):
    stft_kwargs = {
        "frame_len": args.frame_len,
        "frame_hop": args.frame_hop,
        "window": args.window,
        "center": args.center,  # false to comparable with kaldi
        "transpose": True  # T x F
    }
    spectrogram_reader = SpectrogramReader(args.wav_scp, round_power_of_two=args.round_power_of_two,
                                           **stft_kwargs)
    num_done = 0
    with WaveWriter(args.dst_dir, sr=args.sr) as writer:
        for key in spectrogram_reader.index_keys:
            logger.info(f"Processing utt {key}...")
            samps = spectrogram_reader.read(key)
            obs = spectrogram_reader.stft(samps, as_tensor=True)             # N x T x F on the device
            if obs.ndim != 3:
                raise RuntimeError(f"Expected 3D array, but got {obs.ndim}")
            try:
                tf_mask, wpd_enh = facted_wpd(obs, wpd_iters=args.wpd_iters, cgmm_iters=args.cgmm_iters,
                                              update_alpha=args.update_alpha, context=args.context,
                                              taps=args.taps, delay=args.delay)
            except np.linalg.LinAlgError:
                logger.warning(f"{key}: Failed cause LinAlgError in wpd")
                continue
            norm = float(np.max(np.abs(samps)))                              # SpectrogramReader.maxabs
            out = inverse_stft(wpd_enh, norm=norm, **stft_kwargs)
            writer.write(key, out.cpu().numpy())
            if args.dump_mask:
                np.save(f"{args.dst_dir}/{key}", tf_mask[..., 0].cpu().numpy())
            num_done += 1
            if not num_done % 100:
                logger.info(f"Processed {num_done:d} utterances...")
    logger.info(f"Processed {num_done:d} utterances over {len(spectrogram_reader):d}")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Command to do joint dereverbration & denoising algorithm (facted form of WPD)",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Multi-channel rspecifier in kaldi format")
    parser.add_argument("dst_dir", type=str, help="Location to dump enhanced audio")
    parser.add_argument("--taps", default=10, type=int, help="Value of taps used in WPE")
    parser.add_argument("--delay", default=3, type=int, help="Value of delay used in WPE")
    parser.add_argument("--context", default=1, type=int,
                        help="Context value to compute PSD matrix in WPE algorithm")
    parser.add_argument("--wpd-iters", default=3, type=int, help="Number of iterations for WPD")
    parser.add_argument("--cgmm-iters", default=20, type=int, help="Number of iterations for WPD")
    parser.add_argument("--update-alpha", type=strtobool, default=False,
                        help="If true, update alpha in M-step")
    parser.add_argument("--sr", type=int, default=16000, help="Sample rate of the input audio")
    parser.add_argument("--dump-mask", default=False, type=strtobool, help="Dump cgmm mask or not")
    args = parser.parse_args()
    run(args)
