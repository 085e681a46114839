"""Writes deephar_b200/synth_calib/<key>.json: per-BatchNorm-layer (mean, var) scalars and
per-head-conv gains measured by running the CPU oracle (torch fp32) on synthetic frames --
see oracle/synth.py for why.  Run from the repo root:  python tests/golden/make_calibration.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ops_torch, reception, synth  # noqa: E402

OUT = os.path.join('deephar_b200', 'synth_calib')


def calibrate_reception(key, res, **kw):
    cal = synth.Calibrator(1234)
    x = synth.synth_frames(2, res, res, seed=99)
    reception.forward(ops_torch, cal, x, **kw)
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
