"""N > 1 host logic on CPU: world_size-2 and -4 gloo processes shard a clip batch, run a deterministic
stand-in for the per-rank forward, all-gather the outputs, and must reproduce the unsharded
result in the original clip order (also with a batch that does not divide evenly)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deephar_b200.dist import gather_outputs, shard_range


def _fake_forward(x):
    """stand-in for Model.forward_device: one 'action probability' row per clip"""
    return torch.softmax(x.reshape(x.shape[0], 4 * 8 * 8 * 3)[:, :15] * 3.0, dim=-1)


def _worker(rank, world, port, n_clips, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_clips, 4, 8, 8, 3, generator=g)
    a, b = shard_range(n_clips, rank, world)
    local = _fake_forward(x[a:b])
    full = gather_outputs(local, world, n_global=n_clips)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('world,n_clips', [(2, 8), (2, 7), (4, 6), (4, 3)])
def test_shard_and_gather(tmp_path, world, n_clips):
    """even shards, ragged shards (padded to the largest, padding dropped) and a rank with NO clip at all"""
    out = str(tmp_path / 'full.npy')
    mp.spawn(_worker, args=(world, _free_port(), n_clips, out), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n_clips, 4, 8, 8, 3, generator=g)
    ref = _fake_forward(x).numpy()
    got = np.load(out)
    assert got.shape == ref.shape
    assert np.allclose(got, ref, atol=1e-7)


def test_shard_range_covers_everything():
    for n in (1, 7, 16, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, b0), (a1, b1) in zip(spans, spans[1:]):
                assert b0 == a1 and b0 >= a0
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
