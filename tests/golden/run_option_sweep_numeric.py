"""Live numeric check of builder options OFF the BASELINE configs (development container only: needs /root/reference).

For each configuration: the reference's own builder code (unmodified, from /root/reference) is EXECUTED on
tests/golden/keras_shim (eager float64), with the product's synthetic weights assigned by name -- exactly the recipe of
make_reference_golden.py, whose helpers are reused -- and its outputs are compared with the product's COMPILED PLAN for
the same arguments, executed on the CPU by tests/plan_emulator.py.  One JSON line per configuration.

    python tests/golden/run_option_sweep_numeric.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as G  # noqa: E402  (puts the shim, the reference and the repo on sys.path)

sys.path.insert(0, os.path.join(G.ROOT, 'tests'))
import numpy as np  # noqa: E402
from plan_emulator import PlanEmulator  # noqa: E402

from deephar.utils import pose as ref_pose  # noqa: E402
from deephar_b200 import config as pconfig  # noqa: E402


def compare(tag, ref_model, prod_model, x, seed=1234):
    ws = G.collect(ref_model)
    trainable = {n: w for n, w, frozen in ws if not frozen}
    optional = set(prod_model.optional_weights)
    live = {n for n, _ in prod_model.weight_specs if n not in optional}
    assert live == set(trainable), (tag, sorted(live ^ set(trainable))[:6])
    prod_model.init_synthetic_weights(seed)
    table = prod_model.get_weights()
    for n, w in trainable.items():
        w['value'] = np.asarray(table[n], dtype=np.float64)
    ref_outs = ref_model.predict(np.asarray(x, dtype=np.float64))
    ref_outs = ref_outs if isinstance(ref_outs, (list, tuple)) else [ref_outs]
    with np.errstate(over='ignore'):
        outs = PlanEmulator(prod_model).run(x)
    assert len(outs) == len(ref_outs), (tag, len(outs), len(ref_outs))
    err = 0.0
    for o, r in zip(outs, ref_outs):
        r = np.asarray(r, np.float64)
        assert o.shape == r.shape, (tag, o.shape, r.shape)
        err = max(err, float(np.abs(o - r).max() / max(1.0, np.abs(r).max())))
    print(json.dumps({'tag': tag, 'outputs': len(outs), 'launches': len(prod_model.plan.kops), 'max_rel_err': err}), flush=True)


def main():
    rng = np.random.default_rng(2020)
    for kw in (dict(dim=2, num_context_per_joint=None, num_blocks=1, ksize=(3, 3)),
               dict(dim=2, num_context_per_joint=1, num_blocks=2, ksize=(5, 5), export_heatmaps=True),
               dict(dim=2, num_context_per_joint=2, num_blocks=2, ksize=(3, 3), concat_pose_confidence=False, alpha=0.5,
                    export_vfeat_block=1),
               dict(dim=3, depth_maps=8, num_blocks=2, ksize=(3, 3))):
        nj = 17 if kw['dim'] == 3 else 16
        G.fresh_process_state()
        compare('reception %r' % (kw,), G.ref_reception.build((64, 64, 3), nj, **kw), G.reception.build((64, 64, 3), nj, **kw),
                rng.uniform(-1, 1, (2, 64, 64, 3)))
    base = dict(num_pyramids=2, num_levels=4, num_actions=[15], action_pyramids=[1, 2])
    for shape, layout, extra in (
            ((128, 128, 3), 'pa16j2d', dict(num_actions=[], action_pyramids=[], num_levels=3)),
            ((128, 128, 3), 'pa17j3d', dict(num_actions=[], action_pyramids=[], predict_rootz=True, growth=64)),
            ((4, 128, 128, 3), 'pa16j2d', dict(pose_replica=True, kernel_size=(3, 3))),
            ((4, 128, 128, 3), 'pa20j3d', dict(num_actions=[15, 60], sam_alpha=2)),
            ((8, 128, 128, 3), 'pa17j3d', dict(action_pyramids=[2], image_div=4)),
            ((4, 128, 128, 3), 'pa16j2d', dict(num_pyramids=3, action_pyramids=[1, 3], num_pose_features=160,
                                               num_visual_features=96))):
        kw = dict(base)
        kw.update(extra)
        G.fresh_process_state()
        ref = G.ref_spnet.build(G.RefModelConfig(shape, getattr(ref_pose, layout), **kw))
        prod = G.spnet.build(pconfig.ModelConfig(shape, getattr(pconfig, layout), **kw))
        compare('spnet %r %s %r' % (shape, layout, extra), ref, prod, rng.uniform(-1, 1, (1,) + shape))


if __name__ == '__main__':
    main()
