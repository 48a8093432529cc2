#!/usr/bin/env python
"""
tools/ab_fused.py -- time the two fused kernels of config 2 for one library build.

    python tools/ab_fused.py [path/to/libsetk_b200*.so] [label]

Prints one JSON line: ms per call of setk_stft_cov / setk_apply_istft (CUDA events
on the launch stream, 20 calls after 3 warm-ups; inputs exceed L2) and the whole
BeamformPipeline step.  Used to compare builds (ab/*.so) on the same box.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from setk_b200 import _lib, synth  # noqa: E402

lib = sys.argv[1] if len(sys.argv) > 1 else _lib.DEFAULT_LIBRARY
label = sys.argv[2] if len(sys.argv) > 2 else os.path.basename(lib)
_lib.use_library(lib)
from setk_b200.engine import BeamformPipeline  # noqa: E402

B, C, N = int(os.environ.get("AB_BATCH", "256")), 4, 160000
dev = torch.device("cuda:0")
pipe = BeamformPipeline(C, "mvdr", max_batch=B, max_samples=N, device=dev)
a, m = synth.make_batch(8, C, N, device=dev)
audio = a.repeat(B // 8, 1, 1).contiguous()
mask = m.repeat(B // 8, 1, 1).contiguous()


def timed(fn, n=20):
    for _ in range(3):
        r = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r


t_cov, (Rs, Rn, mx) = timed(lambda: pipe.covariances(audio, mask))
t_w, (w, _, _) = timed(lambda: pipe.solve(Rs, Rn))
t_ai, wave = timed(lambda: pipe.plan.apply_istft(audio, w, norm=mx))
t_all, _ = timed(lambda: pipe.run(audio, mask))
chk = float(wave[0].abs().sum()) if torch.is_tensor(wave) else float(wave[0][0].abs().sum())
print(json.dumps({"label": label, "env": os.environ.get("SETK_FUSED_CTAS_PER_SM"),
                  "stft_cov_ms": round(t_cov, 4), "weights_ms": round(t_w, 4),
                  "apply_istft_ms": round(t_ai, 4), "step_ms": round(t_all, 4),
                  "utts_per_s": round(B / t_all * 1e3), "Rs_sum": float(Rs.abs().sum()),
                  "wave_sum": chk}))
