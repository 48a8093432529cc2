"""
Worker of tests/test_gpu_peer_ring.py: one of two processes.  On a one-GPU box they share cuda:0
and exercise the legacy IPC mapping; with two GPUs they also exercise the symmetric-memory mapping
(one GPU per rank).  The rendezvous runs on gloo.
Exits 0 when every batch this pair pushed is on rank 0 bit for bit.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from setk_b200.distributed import PeerResultRing  # noqa: E402


def exercise(mode, dev, rank, world):
    shape = (5, 4001)
    ring = PeerResultRing((shape), torch.int16, dev, slots=3, mode=mode)
    gen = torch.Generator(device="cpu").manual_seed(100 + rank)
    batches = [torch.randint(-32768, 32767, shape, dtype=torch.int16, generator=gen).to(dev) for _ in range(7)]
    for i, b in enumerate(batches):
        ring.push(b * 1, i)             # a temporary: record_stream keeps it alive
    ring.drain()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        for r in range(world):
            g = torch.Generator(device="cpu").manual_seed(100 + r)
            theirs = [torch.randint(-32768, 32767, shape, dtype=torch.int16, generator=g) for _ in range(7)]
            for i in (4, 5, 6):          # the three batches still in the ring
                assert torch.equal(ring.slot(i)[r].cpu(), theirs[i]), (mode, r, i)
    dist.barrier()
    del ring


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the legacy IPC mapping works between two processes on ONE GPU; the symmetric-memory mapping
    # (the fast one: NVLink peer mappings) wants one GPU per rank, so it runs where the box has two
    exercise("ipc", torch.device("cuda:0"), rank, world)
    if torch.cuda.device_count() >= world:
        torch.cuda.set_device(rank)
        exercise("symm", torch.device("cuda", rank), rank, world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
