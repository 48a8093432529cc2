"""
CPU tier: the N > 1 path (utterance sharding + result gather) with world_size 2
on the gloo backend; each rank runs the pipeline on the CPU execution model of
the kernels.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, emu_path, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from setk_b200 import _lib, distributed as D, synth
    from setk_b200.engine import BeamformPipeline
    _lib.use_library(emu_path)
    n_utts, C, N = 5, 2, 3000
    mine = D.shard_indices(n_utts, rank, world)
    pipe = BeamformPipeline(C, "mvdr", max_batch=1, max_samples=N, device="cpu")
    outs = []
    for u in mine:
        x, m = synth.make_batch(1, C, N, device="cpu", first=u)
        wave, status = pipe.run(x, m)
        assert int(status[0]) == 0
        outs.append(wave[0])
    res = D.gather_ragged(outs, dst=0)
    t = D.max_over_ranks(float(rank + 1), torch.device("cpu"))
    assert t == float(world)
    if rank == 0:
        assert sorted(res) == list(range(n_utts))
        np.savez(os.path.join(out_dir, "gathered.npz"), **{str(k): v.numpy() for k, v in res.items()})
    dist.destroy_process_group()


def test_two_rank_shard_and_gather(tmp_path, emu_library_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, emu_library_path, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "gathered.npz")
    # same utterances processed in one process
    sys.path.insert(0, ROOT)
    from setk_b200 import _lib, synth
    from setk_b200.engine import BeamformPipeline
    _lib.use_library(emu_library_path)
    pipe = BeamformPipeline(2, "mvdr", max_batch=1, max_samples=3000, device="cpu")
    for u in range(5):
        x, m = synth.make_batch(1, 2, 3000, device="cpu", first=u)
        wave, _ = pipe.run(x, m)
        assert np.array_equal(got[str(u)], wave[0].numpy())     # bit-identical: deterministic kernels


def test_shard_indices():
    from setk_b200 import distributed as D
    assert D.shard_indices(10, 1, 4) == [1, 5, 9]
    assert sorted(sum((D.shard_indices(7, r, 3) for r in range(3)), [])) == list(range(7))
    assert D.shard_keys(list("abcde"), 0, 2) == ["a", "c", "e"]
