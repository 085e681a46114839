"""End-to-end parity of the SPNet forward (spnet.build) vs the oracle on identical seeded weights
and clips: pose coordinates / confidences / action probabilities <= 1e-3, identical arg-max
action class (north_star)."""
import numpy as np
import pytest

from deephar_b200 import spnet
from deephar_b200.config import ModelConfig, pa16j2d, pa17j3d
from oracle import ops_np, ops_torch, synth
from oracle import spnet as oracle_spnet

pytestmark = pytest.mark.gpu
TOL = 1e-3


def _run(cfg, ocfg, clips, ops, res=128):
    m = spnet.build(cfg).init_synthetic_weights(1234)
    T = cfg.input_shape[0]
    x = np.stack([synth.synth_frames(T, res, res, seed=40 + i) for i in range(clips)])
    refs = oracle_spnet.forward(ops, m.get_weights(), x, ocfg)
    outs = m.predict(x, batch_size=clips)
    assert len(outs) == len(refs)
    for o, r in zip(outs, refs):
        assert o.shape == r.shape, (o.shape, r.shape)
        err = np.abs(o.astype(np.float64) - r).max()
        assert err <= TOL, 'max err %g for output of shape %s' % (err, o.shape)
        if o.ndim == 2:                    # action probabilities: identical arg-max
            assert np.array_equal(o.argmax(-1), r.argmax(-1))
    return m, x, outs


def test_spnet_penn_like_t8(cuda):
    """2-D pose + action, replica heads, T = 8 (time_stride 1) -- fp64 oracle."""
    kw = dict(num_actions=[15], num_pyramids=2, action_pyramids=[1, 2], num_levels=4, pose_replica=True,
              num_pose_features=160, num_visual_features=160)
    cfg = ModelConfig((8, 128, 128, 3), pa16j2d, **kw)
    ocfg = oracle_spnet.ModelConfig((8, 128, 128, 3), oracle_spnet.pa16j2d, **kw)
    m, x, outs = _run(cfg, ocfg, 2, ops_np)
    # split_model + batch-size independence
    pm, am = spnet.split_model(m, cfg)
    a1 = am.predict(x, batch_size=1)
    for a, b in zip(a1, outs[6:]):
        assert np.abs(a - b).max() < 1e-5


def test_spnet_ntu_like_t16(cuda):
    """3-D pose (17 joints -> padded to 20) + 60 actions, T = 16 (time_stride 2)."""
    kw = dict(num_actions=[60], num_pyramids=2, action_pyramids=[1, 2], num_levels=4, num_pose_features=192,
              num_visual_features=192)
    cfg = ModelConfig((16, 128, 128, 3), pa17j3d, **kw)
    ocfg = oracle_spnet.ModelConfig((16, 128, 128, 3), oracle_spnet.pa17j3d, **kw)
    _run(cfg, ocfg, 1, ops_torch)


def test_spnet_pose_only_frames(cuda):
    cfg = ModelConfig((128, 128, 3), pa16j2d, num_pyramids=2, action_pyramids=[], num_levels=4)
    ocfg = oracle_spnet.ModelConfig((128, 128, 3), oracle_spnet.pa16j2d, num_pyramids=2, action_pyramids=[], num_levels=4)
    m = spnet.build(cfg).init_synthetic_weights(5)
    x = synth.synth_frames(3, 128, 128, seed=8)
    refs = oracle_spnet.forward(ops_torch, m.get_weights(), x, ocfg)
    outs = m.predict(x)
    for o, r in zip(outs, refs):
        assert o.shape == r.shape and np.abs(o - r).max() <= TOL
