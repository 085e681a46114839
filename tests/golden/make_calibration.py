"""Writes deephar_b200/synth_calib/<key>.json: per-BatchNorm-layer (mean, var) scalars and
per-head-conv gains measured by running the CPU oracle (torch fp32) on synthetic frames --
see oracle/synth.py for why.  Run from the repo root:  python tests/golden/make_calibration.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import action, ops_torch, reception, spnet, synth  # noqa: E402
import numpy as np  # noqa: E402

OUT = os.path.join('deephar_b200', 'synth_calib')


def calibrate_reception(key, res, **kw):
    cal = synth.Calibrator(1234)
    x = synth.synth_frames(2, res, res, seed=99)
    reception.forward(ops_torch, cal, x, **kw)
    with open(os.path.join(OUT, key + '.json'), 'w') as f:
        json.dump(cal.calib, f, indent=0, sort_keys=True)
    print(key, len(cal.calib), 'entries')


def calibrate_spnet(key, cfg, res, clips):
    cal = synth.Calibrator(1234)
    t = cfg.input_shape[0]
    x = np.stack([synth.synth_frames(t, res, res, seed=90 + i) for i in range(clips)])
    spnet.forward(ops_torch, cal, x, cfg)
    with open(os.path.join(OUT, key + '.json'), 'w') as f:
        json.dump(cal.calib, f, indent=0, sort_keys=True)
    print(key, len(cal.calib), 'entries')


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    calibrate_reception('reception_j16_d2_c2_k5', 128, num_joints=16, dim=2, num_context_per_joint=2,
                        num_blocks=8, ksize=(5, 5))
    calibrate_reception('reception_j17_d3_cNone_k5', 128, num_joints=17, dim=3, num_blocks=8, ksize=(5, 5))
    calibrate_reception('reception_j16_d2_c2_k3', 128, num_joints=16, dim=2, num_context_per_joint=2,
                        num_blocks=8, ksize=(3, 3))
    penn = spnet.ModelConfig((16, 128, 128, 3), spnet.pa16j2d, num_actions=[15], num_pyramids=6,
                             action_pyramids=[5, 6], num_levels=4, pose_replica=True, num_pose_features=160,
                             num_visual_features=160)
    calibrate_spnet('spnet_j16_d2_p6_a5-6_r1_f160', penn, 128, 2)
    ntu = spnet.ModelConfig((16, 128, 128, 3), spnet.pa17j3d, num_actions=[60], num_pyramids=2,
                            action_pyramids=[1, 2], num_levels=4, num_pose_features=192, num_visual_features=192)
    calibrate_spnet('spnet_j17_d3_p2_a1-2_r0_f192', ntu, 128, 2)
    small = spnet.ModelConfig((8, 128, 128, 3), spnet.pa16j2d, num_actions=[15], num_pyramids=2,
                              action_pyramids=[1, 2], num_levels=4, pose_replica=True, num_pose_features=160,
                              num_visual_features=160)
    calibrate_spnet('spnet_j16_d2_p2_a1-2_r1_f160', small, 128, 2)
    cal = synth.Calibrator(1234)
    x = np.stack([synth.synth_frames(16, 128, 128, seed=70 + i) for i in range(2)])
    action.forward(ops_torch, cal, x, 15, 16, 4, num_context_per_joint=2, ksize=(5, 5))
    with open(os.path.join(OUT, 'merge_j16_b4_k5.json'), 'w') as f:
        json.dump(cal.calib, f, indent=0, sort_keys=True)
    print('merge_j16_b4_k5', len(cal.calib), 'entries')
