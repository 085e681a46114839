"""CVPR'18 merge model (action.build_merge_model): host-side weight list vs the oracle (no GPU) and
end-to-end GPU parity of the 9 action outputs."""
import numpy as np
import pytest

from deephar_b200 import action, reception
from oracle import action as oracle_action
from oracle import ops_torch, synth


def _build(res, T, blocks=2):
    pe = reception.build((res, res, 3), 16, dim=2, num_blocks=blocks, num_context_per_joint=2, ksize=(5, 5),
                         concat_pose_confidence=False)
    return action.build_merge_model(pe, 15, (res, res, 3), T, 16, blocks, pose_dim=2)


def test_merge_weight_specs_match_oracle():
    m = _build(64, 16)
    x = synth.synth_frames(16, 64, 64)[None]
    outs, used = oracle_action.forward(ops_torch, synth.SyntheticTable(3), x, 15, 16, 2, 2, (5, 5),
                                       return_weights_used=True)
    assert m.weight_specs == used
    assert len(outs) == 9 and m.output_shape == [(None, 15)] * 9


def test_merge_v2_pose_net_weight_specs_match_oracle():
    """pose_net_version='v2' with 20 frames and the 3-D head (exp/ntu/eval_ntu_ar_pe_merge.py:51-58)."""
    pe = reception.build((64, 64, 3), 20, dim=3, num_blocks=2, depth_maps=8, ksize=(5, 5))
    m = action.build_merge_model(pe, 60, (64, 64, 3), 20, 20, 2, pose_dim=3, depth_maps=8, num_context_per_joint=0,
                                 pose_net_version='v2')
    x = synth.synth_frames(20, 64, 64)[None]
    outs, used = oracle_action.forward(ops_torch, synth.SyntheticTable(3), x, 60, 20, 2, 0, (5, 5), pose_dim=3,
                                       depth_maps=8, pose_net_version='v2', return_weights_used=True)
    assert m.weight_specs == used
    assert any(n.startswith('PoseAR/') and s == (3, 1, 3, 12) for n, s in used)          # 12 filters: the v2 width
    assert len(outs) == 9 and m.output_shape == [(None, 60)] * 9


def test_merge_full_size_config():
    """exp/pennaction/eval_penn_ar_pe_merge.py:42-62: 16 frames, 4 blocks."""
    pe = reception.build((256, 256, 3), 16, dim=2, num_blocks=4, num_context_per_joint=2, ksize=(5, 5),
                         concat_pose_confidence=False)
    m = action.build_merge_model(pe, 15, (256, 256, 3), 16, 16, 4, pose_dim=2)
    assert m.input_shape == (None, 16, 256, 256, 3)
    assert abs(m.conv_flops_per_frame() - 12.0e9) / 12.0e9 < 0.02        # 12.00 GFLOP backbone
    with pytest.raises(ValueError):
        action.build_merge_model(pe, 15, (256, 256, 3), 16, 16, 4, pose_dim=3)   # a 2-D pose network under the 3-D head
    with pytest.raises(ValueError):
        action.build_merge_model(pe, 15, (256, 256, 3), 16, 16, 3, pose_dim=2)   # wrong num_blocks


def test_merge_3d_variant_builds():
    """action.py:208-297 (pose_dim=3): volumetric head; outputs [pose, visible, p1..p4, v1..v4, m]."""
    pe = reception.build((128, 128, 3), 20, dim=3, num_blocks=2, depth_maps=8, ksize=(5, 5))
    m = action.build_merge_model(pe, 60, (128, 128, 3), 8, 20, 2, pose_dim=3, depth_maps=8, output_poses=True)
    assert m.output_shape[:2] == [(None, 8, 20, 3), (None, 8, 20, 1)] and len(m.outputs) == 11
    assert [k.kind for k in m.plan.kops].count('pose_regression_3d_ex') == 1
    with pytest.raises(ValueError):
        action.build_merge_model(pe, 60, (128, 128, 3), 8, 20, 2, pose_dim=3, depth_maps=16)


@pytest.mark.gpu
def test_merge_gpu_parity(cuda):
    m = _build(128, 16).init_synthetic_weights(1234)
    x = np.stack([synth.synth_frames(16, 128, 128, seed=60 + i) for i in range(2)])
    refs = oracle_action.forward(ops_torch, m.get_weights(), x, 15, 16, 2, 2, (5, 5))
    outs = m.predict(x, batch_size=2)
    assert len(outs) == 9
    for o, r in zip(outs, refs):
        assert o.shape == r.shape == (2, 15)
        assert np.abs(o - r).max() <= 1e-3
        assert np.array_equal(o.argmax(-1), r.argmax(-1))
