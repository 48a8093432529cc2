"""
oracle/cpu_bench.py -- time the reference's CPU path (this directory's numpy
restatement of it) on the host cores, for bench.py's `cpu_baseline` and
`--impl reference` legs.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py).  Mirrors how the
reference scales: `run.pl JOB=1:nj` forks one single-threaded Python per
wav.scp shard (scripts/run_adapt_beamformer.sh:66-92); here `workers`
processes each run the per-utterance loop body of
apply_adaptive_beamformer.py:130-177 (forward_stft x C -> MvdrBeamformer.run
-> inverse_stft(norm=max|x|)) on the reference's default dtype path (float32
mask -> complex64 STFT).  No file IO in the timed region.
"""
import multiprocessing as mp
import os
import time

import numpy as np


def _synth(C, N, seed):
    rng = np.random.default_rng(seed)
    env = 0.55 + 0.45 * np.cos(2 * np.pi * 4.0 * np.arange(N) / 16000.0 + rng.uniform(0, 6.28))
    s = rng.standard_normal(N) * env
    taps = np.exp(-np.arange(64) / 8.0)

    def fir(x):
        h = rng.standard_normal(64) * taps
        h[0] = np.sign(h[0]) * (np.abs(h).max() + 0.5)
        return np.convolve(x, h)[:N]

    tgt = np.stack([fir(s) for _ in range(C)])
    noise = np.zeros((C, N))
    for _ in range(C + 2):
        v = rng.standard_normal(N)
        noise += np.stack([fir(v) for _ in range(C)])
    noise += 0.1 * np.std(tgt) * rng.standard_normal((C, N))
    noise *= np.sqrt(np.mean(tgt[0]**2) / (np.mean(noise[0]**2) * 10**0.5))
    mix = tgt + noise
    sc = 0.5 / np.max(np.abs(mix))
    return (mix * sc).astype(np.float32), (tgt[0] * sc).astype(np.float32), \
        (noise[0] * sc).astype(np.float32)


def _worker(args):
    C, N, n_utts, seed = args
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    from oracle import beamformer_oracle as bo
    from oracle import stft_oracle as so
    mix, tgt0, noise0 = _synth(C, N, seed)
    kw = dict(frame_len=512, frame_hop=256, center=True, window="hann", transpose=False)
    S = so.forward_stft(tgt0, round_power_of_two=True, **kw)
    V = so.forward_stft(noise0, round_power_of_two=True, **kw)
    mask = (np.abs(S) / np.sqrt(np.abs(S)**2 + np.abs(V)**2 + so.EPSILON)).T.astype(np.float32)
    bo.enhance_utterance(mix, mask, kind="mvdr", stft_dtype=np.complex64)   # warm-up
    t0 = time.perf_counter()
    for _ in range(n_utts):
        bo.enhance_utterance(mix, mask, kind="mvdr", stft_dtype=np.complex64)
    return time.perf_counter() - t0


def usable_cores():
    """Cores this process may actually use: affinity mask and cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()
        if quota != "max":
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


def throughput(C, N, n_utts_per_worker, workers, seed=20240923):
    """utterances / second over `workers` concurrent single-threaded processes."""
    ctx = mp.get_context("spawn")
    jobs = [(C, N, n_utts_per_worker, seed + i) for i in range(workers)]
    t0 = time.perf_counter()
    with ctx.Pool(workers) as pool:
        elapsed = pool.map(_worker, jobs)
    _ = time.perf_counter() - t0
    return workers * n_utts_per_worker / max(elapsed)
