"""
setk_b200.distributed -- multi-GPU plumbing: one process per GPU, the utterance
list sharded by rank, no collective on the data path, one gather of results at
the end.  This is the B200 equivalent of the reference's
`split_scp.pl` + `run.pl JOB=1:nj` fan-out (scripts/run_adapt_beamformer.sh:66-92):
there every job writes its own files; here rank 0 can collect the enhanced audio
over NCCL (NVLink 5 / NVSwitch) instead.
"""
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


def max_over_ranks(value, device):
    """Max of a Python float over ranks (device-side timing is max over ranks)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
