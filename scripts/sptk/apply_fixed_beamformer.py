#!/usr/bin/env python
# coding=utf-8
"""
Run a fixed beamformer (predefined weights)

Drop-in for the reference's scripts/sptk/apply_fixed_beamformer.py (same positional
arguments, flags, defaults and outputs: <dst_dir>/<key>.wav, PCM-16), with the STFT,
the beamforming and the inverse STFT on libsetk_b200's CUDA kernels (setk_stft,
setk_apply, setk_istft).

Deviation: the reference tests `if beamformer:` (apply_fixed_beamformer.py:41), which is
true for a single FixedBeamformer object too, so F x M weights crash on `beam_index[key]`
with beam_index = None; here the beam index is consulted only for B x F x M weights.
"""
import argparse
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from setk_b200.libs.beamformer import FixedBeamformer  # noqa: E402
from setk_b200.libs.data_handler import ScpReader, SpectrogramReader, WaveWriter  # noqa: E402
from setk_b200.libs.opts import StftParser  # noqa: E402
from setk_b200.libs.utils import get_logger, inverse_stft  # noqa: E402

logger = get_logger(__name__)


def run(args):
    stft_kwargs = {
        "frame_len": args.frame_len,
        "frame_hop": args.frame_hop,
        "window": args.window,
        "center": args.center,
        "transpose": False
    }
    spectrogram_reader = SpectrogramReader(args.wav_scp, round_power_of_two=args.round_power_of_two,
                                           **stft_kwargs)
    weights = np.load(args.weights)                      # F x N or B x F x N
    if weights.ndim == 2:
        beamformer = FixedBeamformer(weights)
        beam_index = None
    else:
        beamformer = [FixedBeamformer(w) for w in weights]
        if not args.beam:
            raise RuntimeError("--beam must be assigned, as there are multiple beams")
        beam_index = ScpReader(args.beam, value_processor=int)
    with WaveWriter(args.dst_dir) as writer:
        for key in spectrogram_reader.index_keys:
            logger.info(f"Processing utterance {key}...")
            samps_in = spectrogram_reader.read(key)
            stft_mat = spectrogram_reader.stft(samps_in, as_tensor=True)       # N x F x T on the device
            if beam_index is not None:
                stft_enh = beamformer[beam_index[key]].run(stft_mat)
            else:
                stft_enh = beamformer.run(stft_mat)
            norm = float(np.max(np.abs(samps_in)))                             # SpectrogramReader.maxabs
            samps = inverse_stft(stft_enh, **stft_kwargs, norm=norm)
            writer.write(key, samps.cpu().numpy() if hasattr(samps, "cpu") else samps)
    logger.info(f"Processed {len(spectrogram_reader):d} utterances")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Command to run fixed beamformer. Runing this command needs "
        "to design fixed beamformer first.",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Multi-channel wave scripts in Kaldi format")
    parser.add_argument("weights", type=str,
                        help="Fixed beamformer weights in numpy format (in shape F x M or B x F x M)")
    parser.add_argument("dst_dir", type=str, help="Location to dump the enhanced audio")
    parser.add_argument("--beam", type=str, default="",
                        help="Beam index to use in beamformer weights (in shape B x F x M)")
    args = parser.parse_args()
    run(args)
