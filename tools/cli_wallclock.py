#!/usr/bin/env python
"""
tools/cli_wallclock.py -- wall-clock throughput of the drop-in command line
scripts/sptk/apply_adaptive_beamformer.py on a synthetic corpus of files (config 2's shape:
4-ch PCM-16 wav, 10 s @ 16 kHz, IRM masks as .npy), written to a RAM disk first.

    python tools/cli_wallclock.py [n_utts=512] [batch_size=256]

Prints one JSON line: utterances per second of the whole process (interpreter start, CUDA
context, file reads, batching, kernels, PCM-16 file writes), and of the feeder alone as the
CLI's own log reports it.  The reference runs the same command one utterance at a time per
process (scripts/run_adapt_beamformer.sh:66-92).
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import scipy.io.wavfile as wavfile  # noqa: E402
import torch  # noqa: E402

from setk_b200 import synth  # noqa: E402
from setk_b200 import plan as P  # noqa: E402


def main():
    n_utts = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="setk_cli_", dir=base)
    try:
        C, N, uniq = 4, 160000, 32
        x, m = synth.make_batch(uniq, C, N, device="cuda")
        pcm = P.float_to_pcm16(x).cpu().numpy()
        m = m.cpu().numpy()
        wav_lines, mask_lines = [], []
        for u in range(n_utts):
            key = f"utt{u:05d}"
            wavfile.write(os.path.join(tmp, key + ".wav"), 16000, np.ascontiguousarray(pcm[u % uniq].T))
            np.save(os.path.join(tmp, key + ".npy"), m[u % uniq])
            wav_lines.append(f"{key} {os.path.join(tmp, key + '.wav')}")
            mask_lines.append(f"{key} {os.path.join(tmp, key + '.npy')}")
        open(os.path.join(tmp, "wav.scp"), "w").write("\n".join(wav_lines) + "\n")
        open(os.path.join(tmp, "mask.scp"), "w").write("\n".join(mask_lines) + "\n")
        del x, m
        torch.cuda.empty_cache()
        dst = os.path.join(tmp, "out")
        cmd = [sys.executable, os.path.join(ROOT, "scripts", "sptk", "apply_adaptive_beamformer.py"),
               "--frame-len", "512", "--frame-hop", "256", "--mask-format", "numpy", "--beamformer", "mvdr",
               "--batch-size", str(batch), "--lookahead", str(2 * batch),
               os.path.join(tmp, "wav.scp"), os.path.join(tmp, "mask.scp"), dst]
        t0 = time.time()
        r = subprocess.run(cmd, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True)
        wall = time.time() - t0
        ok = r.returncode == 0 and len(os.listdir(dst)) == n_utts
        feeder = re.findall(r"Batched feeder: (\d+) utterances.*?([0-9.]+) utts?/s", r.stderr)
        print(json.dumps({"cli": "apply_adaptive_beamformer.py --beamformer mvdr", "n_utts": n_utts,
                          "batch_size": batch, "ok": ok, "wall_s": round(wall, 2),
                          "utts_per_s_whole_process": round(n_utts / wall, 1),
                          "feeder_log": r.stderr.strip().splitlines()[-3:],
                          "feeder_utts_per_s": float(feeder[-1][1]) if feeder else None}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
