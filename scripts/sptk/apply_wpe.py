#!/usr/bin/env python
# coding=utf-8
"""
Do WPE dereverberation

Drop-in for the reference's scripts/sptk/apply_wpe.py (same positional arguments,
flags, defaults and outputs: <dst_dir>/<key>.wav, all channels, PCM-16), with the
STFT, the GWPE iterations and the inverse STFT on libsetk_b200's CUDA kernels
(setk_stft, setk_wpe_stft, setk_istft).  --nara-wpe is accepted and refused: that
package is not part of this build.
"""
import argparse
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import torch  # noqa: E402

from setk_b200.libs.data_handler import WaveReader, WaveWriter  # noqa: E402
from setk_b200.libs.opts import StftParser, strtobool  # noqa: E402
from setk_b200.libs.utils import default_device, get_logger, get_plan  # noqa: E402
from setk_b200.libs.wpe import wpe  # noqa: E402

logger = get_logger(__name__)


def run(args):
    if args.nara_wpe:
        raise RuntimeError("--nara-wpe: the nara_wpe package is not available in this build")
    dev = default_device()
    wave_reader = WaveReader(args.wav_scp, sr=args.sr)
    num_done = 0
    with WaveWriter(args.dst_dir, sr=args.sr) as writer:
        for key, samps in wave_reader:
            logger.info(f"Processing utt {key}...")
            if samps.ndim == 1:
                samps = samps[None]
            C, N = samps.shape
            plan = get_plan(C, args.frame_len, args.frame_hop, bool(args.center),
                            bool(args.round_power_of_two), args.window, N, dev, batch=C)
            audio = torch.from_numpy(np.ascontiguousarray(samps, dtype=np.float32)).to(dev)[None]
            stft = plan.stft(audio)[0]                                   # N x F x T
            try:
                dereverb = wpe(stft.permute(1, 0, 2), num_iters=args.num_iters, context=args.context,
                               taps=args.taps, delay=args.delay)         # F x N x T
            except np.linalg.LinAlgError:
                logger.warning(f"{key}: Failed cause LinAlgError in wpe")
                continue
            # every channel is one "utterance" of the inverse STFT; no `norm` (apply_wpe.py:59-60)
            out = plan.istft(dereverb.permute(1, 0, 2).contiguous())     # N x samples
            writer.write(key, out.cpu().numpy())
            num_done += 1
            if not num_done % 100:
                logger.info(f"Processed {num_done:d} utterances...")
    logger.info(f"Processed {num_done:d} utterances over {len(wave_reader):d}")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Command to do GWPE dereverbration algorithm (recommended "
        "configuration: 512/128/blackman)",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Multi-channel rspecifier in kaldi format")
    parser.add_argument("dst_dir", type=str, help="Location to dump dereverbrated files")
    parser.add_argument("--taps", default=10, type=int, help="Value of taps used in GWPE algorithm")
    parser.add_argument("--delay", default=3, type=int, help="Value of delay used in GWPE algorithm")
    parser.add_argument("--context", default=1, dest="context", type=int,
                        help="Context value to compute PSD matrix in GWPE algorithm")
    parser.add_argument("--num-iters", default=3, type=int, help="Number of iterations to step in GWPE")
    parser.add_argument("--sample-rate", type=int, default=16000, dest="sr",
                        help="Waveform data sample rate")
    parser.add_argument("--nara-wpe", type=strtobool, default=False, help="Use nara-wpe package")
    args = parser.parse_args()
    run(args)
