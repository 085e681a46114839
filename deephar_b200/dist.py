"""Data-parallel plumbing for the forward path: the clip batch shards contiguously
across ranks (one process per GPU), weights replicate, and the ONLY exchange step is an
all-gather of the (small) outputs -- action probabilities (B_local, n_act) and, optionally,
poses.  torch.distributed is used as the transport (NCCL over NVLink on GPUs, gloo in the CPU
tests); there is no collective on the data path itself.
"""


def shard_range(n_items, rank, world):
    """Contiguous split of `n_items` clips/frames: ranks < remainder get one extra item."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_outputs(local, world, group=None):
    """All-gather a per-rank output tensor whose leading axis is the local batch; returns the
    global-batch tensor in rank order (works for unequal shard sizes by padding to the max)."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    n_local = torch.tensor([local.shape[0]], device=local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    nmax = max(sizes)
    if local.shape[0] < nmax:
        pad = torch.zeros((nmax - local.shape[0],) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        local = torch.cat([local, pad], dim=0)
    bufs = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(bufs, local.contiguous(), group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, sizes)], dim=0)
