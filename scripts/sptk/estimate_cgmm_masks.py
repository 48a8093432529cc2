#!/usr/bin/env python
# coding=utf-8
"""
Speech & noise mask estimation with a CGMM

Drop-in for the reference's scripts/sptk/estimate_cgmm_masks.py (same positional
arguments, flags, defaults and outputs: <dst_dir>/<key>.npy, float32, T x F for two
classes, K x T x F otherwise; utterances whose .npy exists are skipped), with the
STFT and the EM iterations on libsetk_b200's CUDA kernels (setk_cgmm_masks: tile
STFT into a bin-major workspace, then fp64 covariance / eigen / posterior kernels;
csrc/cgmm.cu).  Frame sizes other than 512 / 1024 points go through the generic
STFT and setk_cgmm_stft.

Notes against the reference:
  * --num-classes > 2 without --init-mask: the reference starts from
    np.random.uniform posteriors after np.random.seed(--seed); the same numbers
    are drawn here with numpy (once per utterance, in file order) and handed to the
    GPU, so a seeded run follows the reference's stream of random numbers.
  * --num-classes is limited to 4 (the kernels' instantiations).
"""
import argparse
import os
import sys
from pathlib import Path

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import torch  # noqa: E402

from setk_b200.libs.cluster import CgmmTrainer, permu_aligner  # noqa: E402
from setk_b200.libs.data_handler import (NumpyReader, NumpyWriter, ScriptReader,  # noqa: E402
                                         WaveReader)
from setk_b200.libs.opts import StftParser, strtobool  # noqa: E402
from setk_b200.libs.utils import default_device, forward_stft, get_logger, get_plan, nextpow2  # noqa: E402

logger = get_logger(__name__)


def train_utterance(samps, args, init_mask, dev):
    """samps C x N float32 -> masks K x T x F float32 (numpy)."""
    C, N = samps.shape
    n_fft = nextpow2(args.frame_len) if args.round_power_of_two else args.frame_len
    K = args.num_classes
    audio = torch.from_numpy(np.ascontiguousarray(samps)).to(dev)[None]
    init = None
    if init_mask is not None:
        g = torch.from_numpy(np.asarray(init_mask, dtype=np.float32)).to(dev)     # T x F or K x T x F
        if g.dim() == 2:
            g = torch.stack([g, 1 - g])                                            # cluster.py:428-429
        init = g[None].contiguous()
    if n_fft in (512, 1024) and args.frame_hop % 4 == 0 and args.frame_hop <= n_fft:
        plan = get_plan(C, args.frame_len, args.frame_hop, bool(args.center),
                        bool(args.round_power_of_two), args.window, N, dev)
        if init is None and K != 2:
            T, F = plan.num_frames(N), plan.num_bins
            g = np.random.uniform(size=[K, F, T])                                  # cluster.py:432-434
            g = g / np.sum(g, 0, keepdims=True)
            init = torch.from_numpy(np.transpose(g, (0, 2, 1)).astype(np.float32)).to(dev)[None].contiguous()
        masks, status = plan.cgmm_masks(audio, K, args.num_iters, init_gamma=init,
                                        update_alpha=bool(args.update_alpha))
        if int(status.abs().sum()) != 0:
            raise RuntimeError("eigen-iteration did not converge")
        return masks[0].cpu().numpy()
    stft = torch.stack([
        forward_stft(audio[0, c], frame_len=args.frame_len, frame_hop=args.frame_hop,
                     round_power_of_two=bool(args.round_power_of_two), center=bool(args.center),
                     window=args.window, transpose=False) for c in range(C)
    ])                                                                             # C x F x T
    gamma = None
    if init is not None:
        gamma = init[0].transpose(-1, -2)
    elif K != 2:
        g = np.random.uniform(size=[K, stft.shape[1], stft.shape[2]])
        gamma = torch.from_numpy((g / np.sum(g, 0, keepdims=True)).astype(np.float32))
    trainer = CgmmTrainer(stft, K, gamma=gamma, update_alpha=bool(args.update_alpha))
    return trainer.train(args.num_iters).transpose(-1, -2).cpu().numpy()


def run(args):
    np.random.seed(args.seed)
    dev = default_device()
    wave_reader = WaveReader(args.wav_scp, sr=16000)
    MaskReader = {"numpy": NumpyReader, "kaldi": ScriptReader}
    init_mask_reader = MaskReader[args.fmt](args.init_mask) if args.init_mask else None

    num_done = 0
    with NumpyWriter(args.dst_dir) as writer:
        dst_dir = Path(args.dst_dir)
        for key, samps in wave_reader:
            if (dst_dir / f"{key}.npy").exists():
                logger.info(f"Training utterance {key} ... Skip")
                continue
            if samps.ndim == 1:
                samps = samps[None]
            init_mask = None
            if init_mask_reader and key in init_mask_reader:
                init_mask = init_mask_reader[key]          # T x F (or K x T x F)
                logger.info("Using external TF-mask to initialize cgmm")
            try:
                masks = train_utterance(np.asarray(samps, dtype=np.float32), args, init_mask, dev)
                num_done += 1
                if args.solve_permu:
                    masks = permu_aligner(masks)
                    logger.info("Permutation alignment done on each frequency")
                if args.num_classes == 2:
                    masks = masks[0]
                writer.write(key, masks.astype(np.float32))
                logger.info(f"Training utterance {key} ... Done")
            except RuntimeError as err:
                logger.warning(f"Training utterance {key} ... Failed ({err})")
    logger.info(f"Train {num_done:d} utterances over {len(wave_reader):d}")


if __name__ == "__main__":
    parser = argparse.ArgumentParser(
        description="Speech & Noise mask estimation using CGMM model "
        "(also see: estimate_cacgmm_masks.py)",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter,
        parents=[StftParser.parser])
    parser.add_argument("wav_scp", type=str, help="Multi-channel wave scripts in kaldi format")
    parser.add_argument("dst_dir", type=str, help="Location to dump estimated speech masks")
    parser.add_argument("--num-iters", type=int, default=20,
                        help="Number of iterations to train CGMM parameters")
    parser.add_argument("--num-classes", type=int, default=2,
                        help="Number of the cluster used in cacgmm model")
    parser.add_argument("--seed", type=int, default=777, help="Random seed for initialization")
    parser.add_argument("--init-mask", type=str, default="", dest="init_mask",
                        help="Initial TF-mask for cgmm initialization")
    parser.add_argument("--solve-permu", type=strtobool, default=False,
                        help="If true, solving permutation problems")
    parser.add_argument("--update-alpha", type=strtobool, default=False,
                        help="If true, update alpha in M-step")
    parser.add_argument("--mask-format", type=str, dest="fmt", default="numpy",
                        choices=["kaldi", "numpy"], help="Mask storage format")
    args = parser.parse_args()
    run(args)
