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


def gather_outputs(local, world, n_global=None, group=None, out=None, comm=None):
    """All-gather a per-rank output tensor whose leading axis is the local batch; returns the global-batch tensor
    in rank order.  One fixed-shape `all_gather_into_tensor` (NCCL over NVLink / NVSwitch), no host
    synchronisation: the shard sizes follow from `shard_range(n_global, r, world)` on every rank, shards are padded
    to the largest one and the padding is dropped with device-side slicing.  `n_global` defaults to
    world * local_batch (equal shards).  `out`: optional preallocated (world, max_shard, ...) buffer.
    `comm`: a `Comm` (NCCL through the C ABI, dh_allgather_f32); default: torch.distributed."""
    import torch
    import torch.distributed as dist
    if world == 1:
        return local
    n_local = int(local.shape[0])
    if n_global is None:
        n_global = n_local * world
    sizes = [shard_range(n_global, r, world) for r in range(world)]
    sizes = [b - a for a, b in sizes]
    nmax = max(sizes)
    rank = comm.rank if comm is not None else dist.get_rank(group)
    if sizes[rank] != n_local:
        raise ValueError('rank %d holds %d items, shard_range(%d, %d, %d) says %d'
                         % (rank, n_local, n_global, rank, world, sizes[rank]))
    send = local.contiguous()
    if n_local < nmax:
        pad = torch.zeros((nmax - n_local,) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
        send = torch.cat([send, pad], dim=0)
    if out is None:
        out = torch.empty((world, nmax) + tuple(local.shape[1:]), device=local.device, dtype=local.dtype)
    if comm is not None:
        comm.all_gather(send, out=out)
    else:
        dist.all_gather_into_tensor(out.view((world * nmax,) + tuple(local.shape[1:])), send, group=group)
    if all(sz == nmax for sz in sizes):
        return out.view((world * nmax,) + tuple(local.shape[1:]))
    return torch.cat([out[r, :sizes[r]] for r in range(world)], dim=0)


class Comm(object):
    """The exchange step through the C ABI (include/deephar_b200.h: dh_comm_init / dh_allgather_f32, NCCL bound at
    run time inside libdeephar_b200.so).  torch.distributed is only the out-of-band channel that carries rank 0's
    128-byte ncclUniqueId to the other ranks -- any launcher that can move 128 bytes would do."""

    def __init__(self, ctx, rank, world, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _ffi
        self.ctx, self.rank, self.world = ctx, rank, world
        lib = _ffi.lib()
        buf = C.create_string_buffer(128)
        if rank == 0:
            _ffi.check(lib.dh_comm_unique_id(buf), 'dh_comm_unique_id')
        dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
        t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
        dist.broadcast(t, 0, group=group)
        uid = C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)
        _ffi.check(lib.dh_comm_init(ctx.handle, rank, world, uid), 'dh_comm_init')

    def all_gather(self, local, out=None):
        """local: contiguous fp32 CUDA tensor (same shape on every rank) -> (world,) + local.shape, on torch's
        current stream."""
        import torch

        from . import _ffi
        assert local.is_cuda and local.dtype == torch.float32 and local.is_contiguous()
        if out is None:
            out = torch.empty((self.world,) + tuple(local.shape), device=local.device, dtype=torch.float32)
        _ffi.check(_ffi.lib().dh_allgather_f32(self.ctx.handle, local.data_ptr(), out.data_ptr(), local.numel(),
                                               torch.cuda.current_stream().cuda_stream), 'dh_allgather_f32')
        return out
