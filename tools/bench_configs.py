#!/usr/bin/env python
"""
tools/bench_configs.py -- device-resident throughput of the other BASELINE.json
configurations (3: 8-ch GEV / 1024-pt; 4: 6-ch MVDR; 5: {4, 8, 16}-ch sweep).
Informational (the contract line is bench.py's, config 2); prints one JSON
object per configuration with the route each stage took.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from setk_b200 import synth  # noqa: E402
from setk_b200 import plan as P_  # noqa: E402
from setk_b200.engine import BeamformPipeline  # noqa: E402


def run(name, C, frame_len, beamformer, B, N=160000, steps=10, cgmm=False, wpe=False):
    dev = torch.device("cuda:0")
    pipe = BeamformPipeline(C, beamformer, frame_len=frame_len, frame_hop=256, max_batch=B,
                            max_samples=N, device=dev)
    a, m = synth.make_batch(min(B, 4), C, N, device=dev, frame_len=frame_len, n_fft=pipe.plan.n_fft)
    reps = (B + a.shape[0] - 1) // a.shape[0]
    audio = a.repeat(reps, 1, 1)[:B].contiguous()
    mask = m.repeat(reps, 1, 1)[:B].contiguous()
    for _ in range(3):
        wave, status = pipe.run(audio, mask)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        wave, status = pipe.run(audio, mask)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    nf = pipe.plan.n_fft
    route = ("fused" if (C <= 4 and nf == 512) else
             "tile STFT -> bin-major workspace" if nf in (512, 1024) else "generic (explicit STFT in HBM)")

    def timed(fn, n=3):
        fn(); torch.cuda.synchronize()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(n):
            r = fn()
        a1.record(); torch.cuda.synchronize()
        return a0.elapsed_time(a1) / n, r

    t_cov, (Rs, Rn, mx) = timed(lambda: pipe.covariances(audio, mask))
    t_w, (w, _, _) = timed(lambda: pipe.solve(Rs, Rn))
    t_ai, _ = timed(lambda: pipe.plan.apply_istft(audio, w, norm=mx))
    stages = {"stft_cov_ms": t_cov, "weights_ms": t_w, "apply_istft_ms": t_ai}
    cg = None
    if cgmm:
        torch.cuda.synchronize()
        t_cg, (masks, st) = timed(lambda: pipe.plan.cgmm_masks(audio, 2, 20), n=2)
        cg = {"cgmm_20it_ms": t_cg, "cgmm_utts_per_s": B / t_cg * 1e3,
              "mask_mean": float(masks[:, 0].mean()), "status_failures": int((st != 0).sum())}
        stages["cgmm"] = cg
    if wpe:
        S = pipe.plan.stft(audio)
        torch.cuda.synchronize()
        t_w, (out, st) = timed(lambda: P_.wpe_from_stft(S, 10, 3, 1, 3), n=2)
        stages["wpe"] = {"wpe_3it_ms": t_w, "wpe_utts_per_s": B / t_w * 1e3,
                         "status_failures": int((st != 0).sum())}
    print(json.dumps({"config": name, "channels": C, "n_fft": pipe.plan.n_fft, "beamformer": beamformer,
                      "batch": B, "ms_per_batch": ms, "utts_per_s": B / ms * 1e3,
                      "route": route, "stages": stages,
                      "status_failures": int((status != 0).sum())}), flush=True)
    del pipe, audio, mask, wave
    torch.cuda.empty_cache()


CONFIGS = [
    ("cfg2 4ch MVDR 512", 4, 512, "mvdr", 256),
    ("cfg4 6ch MVDR 512", 6, 512, "mvdr", 64),
    ("cfg5 8ch MVDR 512", 8, 512, "mvdr", 64),
    ("cfg5 16ch MVDR 512", 16, 512, "mvdr", 32),
    ("cfg3 8ch GEV 1024", 8, 1024, "gevd", 64),
]

if __name__ == "__main__":
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    for cfg in CONFIGS:
        if only in cfg[0]:
            run(*cfg, steps=steps, cgmm=(len(sys.argv) > 3 and sys.argv[3] == "cgmm"),
                wpe=(len(sys.argv) > 3 and sys.argv[3] == "wpe"))
