"""One rank of a data-parallel forward on the stand-in device (launched by tests/test_host_path.py under
torch.distributed.run + tests/fake_cuda.py --arithmetic): the rank's contiguous shard of a clip batch goes through the
product's Model.forward_device, the per-rank outputs through dist.gather_outputs with a dist.Comm (dh_comm_init /
dh_allgather_f32 of the C ABI -- gloo behind the stand-in), and rank 0 compares the gathered action probabilities with
the oracle's forward of the WHOLE batch.  Prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deephar_b200 import spnet  # noqa: E402
from deephar_b200.config import ModelConfig, pa16j2d  # noqa: E402
from deephar_b200.dist import Comm, gather_outputs, shard_range  # noqa: E402
from oracle import ops_torch, synth  # noqa: E402
from oracle import spnet as oracle_spnet  # noqa: E402


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
    n_clips, T = int(sys.argv[1]), 4
    kw = dict(num_actions=[15], num_pyramids=2, action_pyramids=[1, 2], num_levels=4, num_pose_features=160,
              num_visual_features=160)
    m = spnet.build(ModelConfig((T, 128, 128, 3), pa16j2d, **kw)).init_synthetic_weights(9)
    x = np.stack([synth.synth_frames(T, 128, 128, seed=70 + i) for i in range(n_clips)])
    a, b = shard_range(n_clips, rank, world)
    m._ensure_device_weights()
    comm = Comm(m._ctx, rank, world)
    if b > a:
        outs = m.forward_device(torch.from_numpy(x[a:b]).cuda())
        local = torch.stack([o for o in outs if o.dim() == 2], dim=1)        # (B_local, n_pred, n_act), as bench.py does
        pose = outs[5].reshape(b - a, -1)                                    # last pose output (B_local, T * nj * 3)
    else:
        local, pose = torch.zeros(0, 6, 15), torch.zeros(0, T * 16 * 3)
    full = gather_outputs(local.contiguous(), world, n_global=n_clips, comm=comm)
    full_pose = gather_outputs(pose.contiguous(), world, n_global=n_clips, comm=comm)
    if rank == 0:
        ref = oracle_spnet.forward(ops_torch, m.get_weights(), x, oracle_spnet.ModelConfig((T, 128, 128, 3), oracle_spnet.pa16j2d, **kw))
        ref_pose = np.asarray(ref[5]).reshape(n_clips, -1)
        ref = np.stack([np.asarray(r) for r in ref if np.asarray(r).ndim == 2], axis=1)
        print(json.dumps({'world': world, 'shape': list(full.shape), 'max_err': float(np.abs(full.numpy() - ref).max()),
                          'pose_shape': list(full_pose.shape), 'pose_max_err': float(np.abs(full_pose.numpy() - ref_pose).max()),
                          'argmax_equal': bool(np.array_equal(full.numpy().argmax(-1), ref.argmax(-1)))}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
