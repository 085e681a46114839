"""SPNet host-side checks (no GPU): weight list / order / layouts vs the oracle restatement of
spnet.py, output list, FLOP counts, split_model."""
import numpy as np

from deephar_b200 import spnet
from deephar_b200.config import ModelConfig, pa16j2d, pa17j3d
from oracle import ops_torch, synth
from oracle import spnet as oracle_spnet


def _cfgs(T, res, pyr, act):
    a = ModelConfig((T, res, res, 3), pa16j2d, num_actions=[15], num_pyramids=pyr, action_pyramids=act,
                    num_levels=4, pose_replica=True, num_pose_features=160, num_visual_features=160)
    b = oracle_spnet.ModelConfig((T, res, res, 3), oracle_spnet.pa16j2d, num_actions=[15], num_pyramids=pyr,
                                 action_pyramids=act, num_levels=4, pose_replica=True, num_pose_features=160,
                                 num_visual_features=160)
    return a, b


def test_weight_specs_match_oracle_penn_like():
    cfg, ocfg = _cfgs(8, 128, 2, [1, 2])
    m = spnet.build(cfg)
    x = synth.synth_frames(8, 128, 128)[None]
    _, used = oracle_spnet.forward(ops_torch, synth.SyntheticTable(1), x, ocfg, return_weights_used=True)
    assert m.weight_specs == used


def test_weight_specs_match_oracle_ntu_like():
    cfg = ModelConfig((16, 128, 128, 3), pa17j3d, num_actions=[60], num_pyramids=2, action_pyramids=[2],
                      num_levels=4, num_pose_features=192, num_visual_features=192)
    ocfg = oracle_spnet.ModelConfig((16, 128, 128, 3), oracle_spnet.pa17j3d, num_actions=[60], num_pyramids=2,
                                    action_pyramids=[2], num_levels=4, num_pose_features=192,
                                    num_visual_features=192)
    m = spnet.build(cfg)
    x = synth.synth_frames(16, 128, 128)[None]
    outs, used = oracle_spnet.forward(ops_torch, synth.SyntheticTable(1), x, ocfg, return_weights_used=True)
    assert m.weight_specs == used
    assert [o.shape[1:] for o in outs] == [s[1:] for s in m.output_shape]


def test_c4_config_outputs_and_flops():
    """BASELINE configs[3] (PennAction SPNet, 16-frame clips)."""
    cfg = ModelConfig((16, 256, 256, 3), pa16j2d, num_actions=[15], num_pyramids=6, action_pyramids=[5, 6],
                      num_levels=4, pose_replica=True, num_pose_features=160, num_visual_features=160)
    m = spnet.build(cfg)
    shp = m.output_shape
    assert len(shp) == 18 + 6
    assert shp[:18] == [(None, 16, 16, 3)] * 18 and shp[18:] == [(None, 15)] * 6
    assert m.input_shape == (None, 16, 256, 256, 3)
    per_clip = m.conv_flops_per_frame() * 16
    assert abs(per_clip - 199.07e9) / 199.07e9 < 0.01
    pm, am = spnet.split_model(m, cfg)
    assert len(pm.outputs) == 18 and len(am.outputs) == 6
    assert spnet.get_num_predictions(6, 4) == 18


def test_c5_config_outputs_and_flops():
    """BASELINE configs[4] (NTU SPNet 3-D, 16-frame clips)."""
    cfg = ModelConfig((16, 256, 256, 3), pa17j3d, num_actions=[60], num_pyramids=2, action_pyramids=[1, 2],
                      num_levels=4, num_pose_features=192, num_visual_features=192)
    m = spnet.build(cfg)
    shp = m.output_shape
    assert shp[:6] == [(None, 16, 17, 4)] * 6 and shp[6:] == [(None, 60)] * 6
    per_clip = m.conv_flops_per_frame() * 16
    assert abs(per_clip - 141.46e9) / 141.46e9 < 0.01
    kinds = [k.kind for k in m.plan.kops]
    assert kinds.count('sam2d') == 12 and kinds.count('kron') == 6


def test_pose_only_single_frame_model():
    cfg = ModelConfig((128, 128, 3), pa16j2d, num_pyramids=2, action_pyramids=[], num_levels=4)
    m = spnet.build(cfg)
    assert m.output_shape == [(None, 16, 3)] * 6
    assert 'kron' not in [k.kind for k in m.plan.kops]


def test_optional_weights_are_the_dead_layers_only():
    """Layers the reference builds but keras.Model prunes (they feed no output) are absent from its checkpoints:
    set_weights must accept a table without them and still insist on everything else (spnet.py:249-262)."""
    import numpy as np
    import pytest
    from deephar_b200.weights import synthetic_weights
    cfg = ModelConfig((2, 64, 64, 3), pa16j2d, num_actions=[15], num_pyramids=2, action_pyramids=[1, 2], num_levels=3,
                      pose_replica=True, num_pose_features=32, num_visual_features=32)
    m = spnet.build(cfg)
    assert sorted(m.optional_weights) == ['act4_action_pred_conv3/kernel', 'up2_pb0_heatmaps_conv2/kernel',
                                          'up2_pb0_heatmaps_fw_maps/kernel']
    table = synthetic_weights(m.weight_specs, 1, {})
    for n in m.optional_weights:
        del table[n]
    m.set_weights(table)                                   # checkpoint-like table: accepted
    assert all(np.all(m.get_weights()[n] == 0) for n in m.optional_weights)
    required = next(n for n, _ in m.weight_specs if n not in m.optional_weights)
    del table[required]
    with pytest.raises(KeyError):
        m.set_weights(table)
