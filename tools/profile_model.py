"""Per-op CUDA-event profile of one forward (micro-batch N) of a model: label, ms, launches,
algorithmic TFLOP/s for convs.  usage: python tools/profile_model.py reception2d|spnet_penn|spnet_ntu [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else 'reception2d'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
if wl in ('reception2d', 'reception3d'):
    from deephar_b200 import reception
    kw = bench.MODEL_KW if wl == 'reception2d' else dict(num_joints=17, dim=3, num_blocks=8, ksize=(5, 5),
                                                          concat_pose_confidence=False)
    m = reception.build((256, 256, 3), **kw).init_synthetic_weights(1234)
    x = torch.rand(n, 256, 256, 3, device='cuda') * 2 - 1
else:
    from deephar_b200 import spnet
    from deephar_b200.config import ModelConfig, pa16j2d, pa17j3d
    if wl == 'spnet_penn':
        cfg = ModelConfig((16, 256, 256, 3), pa16j2d, num_actions=[15], num_pyramids=6, action_pyramids=[5, 6],
                          num_levels=4, pose_replica=True, num_pose_features=160, num_visual_features=160)
    else:
        cfg = ModelConfig((16, 256, 256, 3), pa17j3d, num_actions=[60], num_pyramids=2, action_pyramids=[1, 2],
                          num_levels=4, num_pose_features=192, num_visual_features=192)
    m = spnet.build(cfg).init_synthetic_weights(1234)
    x = torch.rand(n // 16, 16, 256, 256, 3, device='cuda') * 2 - 1
for _ in range(2):
    m.forward_device(x)
torch.cuda.synchronize()
prof = m.profile(x)
tot = sum(r['ms'] for r in prof.values())
print('%s  N=%d  total %.2f ms  (%d ops)' % (wl, n, tot, sum(r['launches'] for r in prof.values())))
for r in sorted(prof.values(), key=lambda r: -r['ms']):
    tf = r['flops'] / (r['ms'] / r['launches']) / 1e9 if r['flops'] else 0.0
    print('%-58s %8.3f ms %5.1f%%  x%-3d %s' % (r['label'], r['ms'], 100 * r['ms'] / tot, r['launches'],
                                                  ('%.0f TFLOP/s' % tf) if tf else ''))
