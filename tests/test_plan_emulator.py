"""The compiled plans, executed on the CPU (tests/plan_emulator.py): every model family's launch sequence -- fused conv
prologues / epilogues, upsampled residuals, pooled second outputs, concat views and copies, fused heads, liveness-aliased
buffers -- run op by op on the plan's own buffer layout in float64 equals the oracle's layer-by-layer forward to 1e-10.
No GPU: this pins compiler.py + the planner, not the kernels."""
import numpy as np
import pytest

from deephar_b200 import action, reception, spnet
from deephar_b200.config import ModelConfig, pa16j2d, pa17j3d
from deephar_b200.weights import fold_batchnorm
from oracle import action as oracle_action
from oracle import ops_np, synth
from oracle import reception as oracle_reception
from oracle import spnet as oracle_spnet
from plan_emulator import PlanEmulator

TOL = 1e-10


def _same(outs, refs):
    assert len(outs) == len(refs)
    for i, (o, r) in enumerate(zip(outs, refs)):
        r = np.asarray(r, np.float64)
        assert o.shape == r.shape, (i, o.shape, r.shape)
        assert np.isfinite(o).all(), 'output %d holds NaN: a launch read memory nothing had written' % i
        assert np.abs(o - r).max() <= TOL * max(1.0, np.abs(r).max()), (i, float(np.abs(o - r).max()))


@pytest.mark.parametrize('kw', [
    dict(num_joints=16, dim=2, num_context_per_joint=2, num_blocks=2, ksize=(5, 5), concat_pose_confidence=False),
    dict(num_joints=16, dim=2, num_context_per_joint=2, num_blocks=2, ksize=(3, 3), export_heatmaps=True),
    dict(num_joints=16, dim=2, num_context_per_joint=None, num_blocks=1, ksize=(5, 5), concat_pose_confidence=False),
    dict(num_joints=17, dim=3, num_blocks=2, ksize=(5, 5), concat_pose_confidence=False),
    dict(num_joints=17, dim=3, depth_maps=8, num_blocks=1, ksize=(3, 3)),
], ids=['2d_ctx', '2d_ctx_concat_heatmaps', '2d_plain', '3d', '3d_concat_d8'])
def test_reception_plans(kw):
    m = reception.build((64, 64, 3), **kw).init_synthetic_weights(3)
    x = synth.synth_frames(3, 64, 64, seed=1).astype(np.float64)
    emu = PlanEmulator(m)
    _same(emu.run(x), oracle_reception.forward(ops_np, m.get_weights(), x, **kw))
    assert emu.launches == len(m.plan.kops)


def test_reception_plan_at_the_headline_geometry():
    """256 x 256 input, 5x5 kernels: 32-pixel-wide maps, where the upsampled residual goes into the separable conv's
    epilogue and the block-end max-pool is the fReMap kernel's second output (2 blocks keep the CPU time down)."""
    kw = dict(num_joints=16, dim=2, num_context_per_joint=2, num_blocks=2, ksize=(5, 5), concat_pose_confidence=False)
    m = reception.build((256, 256, 3), **kw).init_synthetic_weights(1234)
    kinds = [k for k in m.plan.kops]
    assert any(k.attrs.get('res_up2x') for k in kinds if k.kind == 'sepconv')
    assert any(k.attrs.get('pool_out') for k in kinds if k.kind == 'conv')
    x = synth.synth_frames(1, 256, 256, seed=4).astype(np.float64)
    _same(PlanEmulator(m).run(x), oracle_reception.forward(ops_np, m.get_weights(), x, **kw))


@pytest.mark.parametrize('T,layout,olayout,kw', [
    (4, pa16j2d, oracle_spnet.pa16j2d, dict(num_actions=[15], num_pyramids=2, action_pyramids=[1, 2], num_levels=4,
                                            pose_replica=True, num_pose_features=160, num_visual_features=160)),
    (16, pa17j3d, oracle_spnet.pa17j3d, dict(num_actions=[60], num_pyramids=2, action_pyramids=[1, 2], num_levels=4,
                                             num_pose_features=192, num_visual_features=192)),
    (None, pa16j2d, oracle_spnet.pa16j2d, dict(num_pyramids=2, action_pyramids=[], num_levels=4)),
], ids=['penn_like_t4_replica', 'ntu_like_t16_3d', 'pose_only_frames'])
def test_spnet_plans(T, layout, olayout, kw):
    shape = (128, 128, 3) if T is None else (T, 128, 128, 3)
    m = spnet.build(ModelConfig(shape, layout, **kw)).init_synthetic_weights(5)
    ocfg = oracle_spnet.ModelConfig(shape, olayout, **kw)
    x = synth.synth_frames(T or 2, 128, 128, seed=2).astype(np.float64)
    x = x if T is None else x[None]
    with np.errstate(over='ignore'):
        _same(PlanEmulator(m).run(x), oracle_spnet.forward(ops_np, m.get_weights(), x, ocfg))


@pytest.mark.parametrize('pose_dim', [2, 3])
def test_merge_model_plans(pose_dim):
    if pose_dim == 2:
        pe = reception.build((64, 64, 3), 16, dim=2, num_blocks=2, num_context_per_joint=2, ksize=(5, 5),
                             concat_pose_confidence=False)
        m = action.build_merge_model(pe, 15, (64, 64, 3), 16, 16, 2, pose_dim=2)
        args = (15, 16, 2, 2, (5, 5))
        okw = {}
    else:
        pe = reception.build((64, 64, 3), 20, dim=3, num_blocks=2, depth_maps=8, ksize=(5, 5))
        m = action.build_merge_model(pe, 60, (64, 64, 3), 20, 20, 2, pose_dim=3, depth_maps=8, num_context_per_joint=0,
                                     pose_net_version='v2', output_poses=True)
        args = (60, 20, 2, 0, (5, 5))
        okw = dict(pose_dim=3, depth_maps=8, pose_net_version='v2', output_poses=True)
    m.init_synthetic_weights(11)
    T = m.graph.frames_per_clip
    x = synth.synth_frames(T, 64, 64, seed=3).astype(np.float64)[None]
    outs = PlanEmulator(m).run(x)
    refs = oracle_action.forward(ops_np, m.get_weights(), x, *args, **okw)
    _same(outs, refs)
    assert all(abs(float(o.sum()) - 1.0) < 1e-9 for o in outs[-9:])          # nine action distributions


def test_keras_style_model_plan():
    """a model recorded from Keras-style code (keras_compat): strided first conv, BN without gamma, separable residual,
    TimeDistributed layer, upsample-add, channel soft-max head, concatenated feature output"""
    from test_keras_compat import _models, _oracle_forward
    mk, _ = _models()
    mk.init_synthetic_weights(7)
    x = np.random.default_rng(0).uniform(-1, 1, (3, 32, 32, 3))
    outs = PlanEmulator(mk).run(x)
    refs = _oracle_forward(mk.get_weights(), x.astype(np.float32))               # torch fp32 oracle
    for o, r in zip(outs, refs):
        assert o.shape == r.shape and np.abs(o - r).max() <= 1e-4


def test_emulator_catches_a_broken_plan():
    """the emulator is only a check if it fails on a wrong plan: (1) a producer scheduled after its consumer -> the
    consumer reads NaN-filled memory; (2) a dropped residual -> finite but different outputs"""
    kw = dict(num_joints=16, dim=2, num_context_per_joint=2, num_blocks=1, ksize=(3, 3), concat_pose_confidence=False)
    m = reception.build((64, 64, 3), **kw).init_synthetic_weights(3)
    x = synth.synth_frames(1, 64, 64, seed=1).astype(np.float64)
    refs = oracle_reception.forward(ops_np, m.get_weights(), x, **kw)
    kops = m.plan.kops
    try:
        m.plan.kops = [kops[1], kops[0]] + kops[2:]
        assert not np.isfinite(PlanEmulator(m).run(x)[0]).all()
        m.plan.kops = kops
        victim = next(k for k in kops if k.kind == 'sepconv' and k.attrs['n_res'] >= 1)
        victim.attrs['n_res'] -= 1
        out = PlanEmulator(m).run(x)[0]
        victim.attrs['n_res'] += 1
        assert np.isfinite(out).all() and np.abs(out - refs[0]).max() > 1e-6
    finally:
        m.plan.kops = kops
    _same(PlanEmulator(m).run(x), refs)


def test_fp32_batchnorm_fold_of_the_product_equals_the_float64_fold():
    rng = np.random.default_rng(0)
    gamma, beta, mean = (rng.uniform(0.8, 1.2, 64).astype(np.float32), rng.normal(0, 0.1, 64).astype(np.float32),
                         rng.normal(0, 0.1, 64).astype(np.float32))
    var = rng.uniform(0.5, 1.5, 64).astype(np.float32)
    for g in (gamma, None):
        scale, shift = fold_batchnorm(g, beta, mean, var)
        want = (1.0 if g is None else g.astype(np.float64)) / np.sqrt(var.astype(np.float64) + 1e-3)
        assert scale.dtype == shift.dtype == np.float32
        assert np.abs(scale - want).max() <= 1e-7 * np.abs(want).max()
        assert np.abs(shift - (beta - mean.astype(np.float64) * want)).max() <= 2e-7
