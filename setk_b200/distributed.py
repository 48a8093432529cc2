"""
setk_b200.distributed -- multi-GPU plumbing: one process per GPU, the utterance
list sharded by rank, no collective on the data path, results collected on rank 0
 -- by peer copies over NVLink (`PeerResultRing`) or NCCL.  This is the B200 equivalent of the reference's
`split_scp.pl` + `run.pl JOB=1:nj` fan-out (scripts/run_adapt_beamformer.sh:66-92):
there every job writes its own files; here rank 0 can collect the enhanced audio
over NCCL (NVLink 5 / NVSwitch) instead.
"""
import os

import torch
import torch.distributed as dist


def shard_indices(num_items, rank, world_size):
    """Round-robin shard, like split_scp.pl's default: item i -> rank i % W."""
    return list(range(rank, num_items, world_size))


def shard_keys(keys, rank, world_size):
    return [keys[i] for i in shard_indices(len(keys), rank, world_size)]


def gather_results(local, dst=0, group=None):
    """
    Gather equally shaped per-rank result tensors to `dst`.
    Returns a list of W tensors on dst, None elsewhere.
    """
    if not dist.is_available() or not dist.is_initialized():
        return [local]
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    out = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    dist.gather(local, out, dst=dst, group=group)
    return out


def gather_ragged(local_list, dst=0, group=None):
    """
    Gather per-utterance 1-D tensors of different lengths (each rank may hold a
    different number): lengths first, then one padded payload.
    Returns {global_index: tensor} on dst for the round-robin sharding above.
    """
    if not dist.is_available() or not dist.is_initialized():
        return {i: t for i, t in enumerate(local_list)}
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = local_list[0].device if local_list else torch.device("cpu")
    n_local = torch.tensor([len(local_list)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local, group=group)
    max_n = int(max(int(c) for c in counts))
    lens = torch.zeros((max_n,), dtype=torch.int64, device=dev)
    for i, t in enumerate(local_list):
        lens[i] = t.numel()
    all_lens = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_lens, lens, group=group)
    max_len = int(max(int(l.max()) for l in all_lens)) if max_n else 0
    dtype = local_list[0].dtype if local_list else torch.float32
    payload = torch.zeros((max_n, max_len), dtype=dtype, device=dev)
    for i, t in enumerate(local_list):
        payload[i, :t.numel()] = t
    out = [torch.empty_like(payload) for _ in range(world)] if rank == dst else None
    dist.gather(payload, out, dst=dst, group=group)
    if rank != dst:
        return None
    result = {}
    for r in range(world):
        for i in range(int(counts[r])):
            result[r + i * world] = out[r][i, :int(all_lens[r][i])].clone()
    return result


class PeerResultRing:
    """
    Result collection over NVLink peer memory instead of a collective kernel.

    Rank `dst` owns a ring of `slots` receive buffers, each `[world, *shape]`; every other rank
    maps that allocation into its own address space (CUDA IPC handle, exchanged once through the
    process group) and delivers a batch with ONE asynchronous device-to-peer copy on a side stream:
    the copy engines move the bytes over NVLink / NVSwitch while the SMs of both GPUs keep running
    the next batch.  `ncclGather` needs a kernel on every rank for the same transfer; the fused
    kernels of this path are persistent (one wave of 2 x 148 CTAs) and lose their balance when a
    collective kernel holds an SM, which is why the bench measured 0.89 weak-scaling efficiency at
    two GPUs with the NCCL gather of every batch.

    The reference has no counterpart (its jobs write their own files,
    scripts/run_adapt_beamformer.sh:66-92); SURVEY.md section 8e defines the collective as "all
    results to rank 0".  Flow control is the consumer's business: slot `i % slots` is overwritten
    `slots` batches later, so whoever reads the ring on `dst` must be done by then (the bench only
    checks the last batch).
    """

    def __init__(self, shape, dtype, device, slots=3, dst=0, group=None, mode=None):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.dst = dst
        self.slots = slots
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._owned = self._hdl = None
        full = (slots, self.world) + tuple(shape)
        mode = mode or os.environ.get("SETK_PEER_RING", "symm")
        if mode == "symm":
            # torch symmetric memory: every rank allocates the ring with the driver's virtual-memory
            # API and maps its peers' copies (POSIX handles over the group's store) -- real NVLink
            # peer mappings.  Only dst's copy is ever written.
            import torch.distributed._symmetric_memory as symm
            self._owned = symm.empty(full, dtype=dtype, device=self.device)
            self._owned.zero_()
            torch.cuda.synchronize(self.device)
            self._hdl = symm.rendezvous(self._owned, group if group is not None else dist.group.WORLD)
            self.ring = self._owned if self.rank == dst else self._hdl.get_buffer(dst, full, dtype)
        else:
            # legacy CUDA IPC handle of a caching-allocator block (measured 35 GB/s between two B200s
            # of an NVSwitch box: the mapping does not ride NVLink there; kept as a functional fallback)
            from torch.multiprocessing.reductions import reduce_tensor
            payload = [None]
            if self.rank == dst:
                self._owned = torch.zeros(full, dtype=dtype, device=self.device)
                torch.cuda.synchronize(self.device)
                payload = [reduce_tensor(self._owned)]
            dist.broadcast_object_list(payload, src=dst, group=group)
            if self.rank == dst:
                self.ring = self._owned
            else:
                rebuild, rebuild_args = payload[0]
                self.ring = rebuild(*rebuild_args)      # a tensor on dst's GPU, mapped in this process
        self.mode = mode
        self.mine = [self.ring[s, self.rank] for s in range(slots)]

    @classmethod
    def create(cls, shape, dtype, device, slots=3, dst=0, group=None, min_gbps=100.0):
        """
        The ring, or None on EVERY rank when any rank cannot map the peer allocation or its copies
        run below `min_gbps` (a mapping that does not ride NVLink is worse than the collective).
        Tries the symmetric-memory mapping, then the legacy IPC handle.
        """
        world = dist.get_world_size(group)

        def agree(ok):                       # the same decision on every rank
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return int(flag.item()) == 1

        for mode in ("symm", "ipc"):
            ring = None
            try:
                ring = cls(shape, dtype, device, slots, dst, group, mode=mode)
                probe = torch.zeros(tuple(shape), dtype=dtype, device=device)
                ring.push(probe, 0)
                ring.drain()
                torch.cuda.synchronize(device)
                ok = True
            except Exception as err:     # noqa: BLE001 -- any failure means "try the next transport"
                ok = False
                cls.last_error = f"{mode}: {err!r}"
            if not agree(ok):
                ring = None
                continue
            # every rank holds a working mapping: time three deliveries
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier(group)
            e0.record()
            for i in range(3):
                ring.push(probe, i)
            ring.drain()
            e1.record()
            torch.cuda.synchronize(device)
            nbytes = probe.numel() * probe.element_size()
            gbps = 3 * nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9
            cls.last_gbps = gbps
            fast = ring.rank == dst or world == 1 or nbytes < (1 << 24) or gbps >= min_gbps
            if not fast:
                cls.last_error = f"{mode}: peer copies run at {gbps:.0f} GB/s (< {min_gbps:.0f})"
            if agree(fast):
                return ring
            ring = None
        return None

    last_error = None
    last_gbps = None

    def push(self, result, step):
        """Deliver this rank's `result` into slot `step % slots` (asynchronous, ordered after the
        current stream's work so far)."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.mine[step % self.slots].copy_(result, non_blocking=True)
        result.record_stream(self.stream)

    def drain(self):
        """Order the current stream after every copy pushed so far."""
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def slot(self, step):
        """[world, *shape] view of the slot batch `step` landed in (meaningful on dst)."""
        return self.ring[step % self.slots]


def max_over_ranks(value, device):
    """Max of a Python float over ranks (device-side timing is max over ranks)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
