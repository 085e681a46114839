"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN MODEL BUILDERS (runs only where
/root/reference exists; the fixtures it writes are what the tests read).

    python tests/golden/make_reference_golden.py            # writes tests/golden/ref_*.npz

keras / tensorflow are not installable here, so the reference code (deephar/models/reception.py, spnet.py,
action.py, blocks.py, common.py, layers.py, activations.py -- imported unmodified from /root/reference) runs on
tests/golden/keras_shim: an eager float64 stand-in for the Keras 2.1.4 functional API (README there).  For
every case the script
  1. builds the reference model and lists its weights as Keras would save them
     ("<sub-model>/<layer>/<weight>", auto-names from Keras's per-class counters), separating the constants
     the reference code itself assigns with set_weights (soft-argmax grids, aggregation matrix);
  2. builds the product model (deephar_b200) for the same arguments, checks that its weight_specs minus
     `optional_weights` are exactly the reference's trainable weights (names AND shapes), fills the reference
     model with the product's synthetic weights BY NAME;
  3. runs the reference graph on seeded inputs and stores inputs' seed, outputs and the weight list.
tests/test_reference_golden.py then checks the oracle (both op sets) and, on the GPU, the product against these
files.  What is and is not pinned by this: tests/golden/keras_shim/README.md.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get('DEEPHAR_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'keras_shim'))
sys.path.insert(1, REFERENCE)
sys.path.insert(2, ROOT)
sys.path.insert(3, HERE)
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402

np.seterr(all='ignore')

import keras  # noqa: E402,F401  (the shim)
from keras.engine import Model as KModel, reset_uids  # noqa: E402
from keras.layers import TimeDistributed  # noqa: E402

assert keras.__version__.endswith('shim')
import deephar  # noqa: E402,F401  (the reference, unmodified)
from deephar.config import ModelConfig as RefModelConfig  # noqa: E402
from deephar.models import action as ref_action  # noqa: E402
from deephar.models import reception as ref_reception  # noqa: E402
from deephar.models import spnet as ref_spnet  # noqa: E402
from deephar.utils.pose import pa16j2d as ref_pa16j2d, pa17j3d as ref_pa17j3d  # noqa: E402

from deephar_b200 import action, reception, spnet  # noqa: E402
from deephar_b200.config import ModelConfig, pa16j2d, pa17j3d  # noqa: E402
from oracle import synth  # noqa: E402


def fresh_process_state():
    """Every reference script builds ONE model per process.  Emulate that between cases: reset Keras's
    auto-name counters and the module-level counters the reference keeps in globals()
    (spnet.py:210 act_cnt, layers.py:167 global_sam_cnt, :412 max_min_pool_cnt, :429 global_max_min_pool_cnt)."""
    import deephar.layers as ref_layers
    reset_uids()
    ref_spnet.__dict__.pop('act_cnt', None)
    for name in ('global_sam_cnt', 'max_min_pool_cnt', 'global_max_min_pool_cnt'):
        ref_layers.__dict__.pop(name, None)


def collect(model, prefix='', trainable=True, strip=()):
    """Weights as keras.Model.save_weights groups them: nested models contribute '<model>/<layer>/<weight>'.
    Returns [(name, weight, frozen)]: frozen = a constant the reference code assigned with set_weights on a layer
    (or inside a model) it then marked non-trainable (soft-argmax grids, aggregation matrix) -- never in a
    checkpoint's learned state.  `strip`: wrapper prefixes to drop (the merge model wraps the pose network's
    sub-models in TimeDistributed layers named 'td_<name>', action.py:117-153)."""
    out = []
    for l in model.layers:
        inner = l.layer if isinstance(l, TimeDistributed) else l
        eff = trainable and getattr(l, 'trainable', True) and getattr(inner, 'trainable', True)
        if isinstance(inner, KModel):
            name = l.name
            for pre in strip:
                if name.startswith(pre):
                    name = name[len(pre):]
            sub = '' if name.startswith('Model') else name + '/'          # anonymous 'Model<N>' wrapper: transparent
            out += collect(inner, prefix + sub, eff and getattr(inner, 'trainable', True), strip)
        else:
            for w in l.weights:
                out.append((prefix + l.name + '/' + w['name'], w, w['fixed'] and not eff))
    return out


def pin(case, ref_model, prod_model, seed, x, strip=(), x_recipe=None, weight_hook=None):
    ws = collect(ref_model, strip=strip)
    names = [n for n, _, _ in ws]
    assert len(set(names)) == len(names), 'duplicate weight names in the reference model'
    trainable = {n: w for n, w, frozen in ws if not frozen}
    specs = dict(prod_model.weight_specs)
    live = {n: s for n, s in specs.items() if n not in set(prod_model.optional_weights)}
    assert set(live) == set(trainable), (sorted(set(live) - set(trainable))[:5], sorted(set(trainable) - set(live))[:5])
    for n, w in trainable.items():
        assert tuple(w['value'].shape) == tuple(specs[n]), (n, w['value'].shape, specs[n])
    prod_model.init_synthetic_weights(seed)
    table = prod_model.get_weights()
    if weight_hook is not None:
        table = weight_hook(table)
    for n, w in trainable.items():
        w['value'] = np.asarray(table[n], dtype=np.float64)
    outs = ref_model.predict(np.asarray(x, dtype=np.float64))
    if not isinstance(outs, (list, tuple)):
        outs = [outs]
    path = os.path.join(HERE, 'ref_%s.npz' % case)
    # big outputs (exported heat-maps) are kept in float32: they are compared at 1e-3 / by arg-max
    blob = {'out%d' % i: np.asarray(o, dtype=np.float64 if np.size(o) < 10000 else np.float32) for i, o in enumerate(outs)}
    blob['weight_names'] = np.array(sorted(trainable))
    blob['weight_shapes'] = np.array([repr(tuple(trainable[n]['value'].shape)) for n in sorted(trainable)])
    blob['fixed_names'] = np.array(sorted(n for n, w, frozen in ws if frozen))
    blob['optional_in_product'] = np.array(sorted(prod_model.optional_weights))
    blob['seed'] = np.array(seed)
    if x_recipe is None:
        blob['x'] = np.asarray(x, dtype=np.float32)
    else:           # large inputs: the test regenerates them; keep the recipe and a checksum
        import zlib
        blob['x_recipe'] = np.array(x_recipe)
        blob['x_crc32'] = np.array(zlib.crc32(np.ascontiguousarray(x, dtype=np.float32).tobytes()))
    np.savez_compressed(path, **blob)
    print('%-28s %3d trainable + %2d fixed weights, %d outputs -> %s (%.0f KB)' % (
        case, len(trainable), len(ws) - len(trainable), len(outs), os.path.basename(path), os.path.getsize(path) / 1024.0))


def main():
    from ref_cases import LARGE_INPUT_CASES, MERGE3D_CASE, MERGE_CASE, RECEPTION_CASES, SPNET_CASES, SPNET_FULL_CASES
    only = set(sys.argv[1:])
    # ---- ReceptionNet (CVPR'18) ----
    for case, spec in RECEPTION_CASES.items():
        shape, kw, seed, xs = spec[:4]
        frames = spec[4] if len(spec) > 4 else 2
        if only and case not in only:
            continue
        fresh_process_state()
        ref = ref_reception.build(shape, **kw)
        prod = reception.build(shape, **kw)
        pin(case, ref, prod, seed, synth.synth_frames(frames, shape[0], shape[1], seed=xs),
            x_recipe='synth_frames(%d,%d,%d,seed=%d)' % (frames, shape[0], shape[1], xs) if case in LARGE_INPUT_CASES else None)

    # ---- SPNet (TPAMI'20) ----
    rng = np.random.default_rng(11)
    layouts = {'pa16j2d': (ref_pa16j2d, pa16j2d), 'pa17j3d': (ref_pa17j3d, pa17j3d)}
    for case, (shape, layout, kw, seed, batch) in SPNET_CASES.items():
        x_case = rng.uniform(-1.0, 1.0, (batch,) + shape)          # always drawn: keeps the shared stream's order
        if only and case not in only:
            continue
        fresh_process_state()
        ref = ref_spnet.build(RefModelConfig(shape, layouts[layout][0], **kw))
        prod = spnet.build(ModelConfig(shape, layouts[layout][1], **kw))
        pin(case, ref, prod, seed, x_case)

    for case, (shape, layout, kw, seed, batch, xs) in SPNET_FULL_CASES.items():
        if only and case not in only:
            continue
        fresh_process_state()
        ref = ref_spnet.build(RefModelConfig(shape, layouts[layout][0], **kw))
        prod = spnet.build(ModelConfig(shape, layouts[layout][1], **kw))
        pin(case, ref, prod, seed, np.random.default_rng(xs).uniform(-1.0, 1.0, (batch,) + shape),
            x_recipe='default_rng(%d).uniform(-1,1,%r)' % (xs, (batch,) + shape))

    # ---- CVPR'18 merge model (2-D pose + action) ----
    x_merge = rng.uniform(-1.0, 1.0, (1, MERGE_CASE['num_frames']) + MERGE_CASE['input_shape'])
    if not only or 'merge3d_model' in only:
        fresh_process_state()
        mc = MERGE3D_CASE
        ref_pe = ref_reception.build(mc['input_shape'], **mc['reception'])
        ref = ref_action.build_merge_model(ref_pe, mc['num_actions'], mc['input_shape'], mc['num_frames'], mc['num_joints'],
                                           mc['num_blocks'], pose_dim=3, depth_maps=mc['depth_maps'], output_poses=True)
        prod_pe = reception.build(mc['input_shape'], **mc['reception'])
        prod = action.build_merge_model(prod_pe, mc['num_actions'], mc['input_shape'], mc['num_frames'], mc['num_joints'],
                                        mc['num_blocks'], pose_dim=3, depth_maps=mc['depth_maps'], output_poses=True)
        pin('merge3d_model', ref, prod, mc['seed'],
            np.random.default_rng(77).uniform(-1.0, 1.0, (2, mc['num_frames']) + mc['input_shape']), strip=('td_',))
    if only and 'merge_model' not in only:
        return
    fresh_process_state()
    mc = MERGE_CASE
    ref_pe = ref_reception.build(mc['input_shape'], **mc['reception'])
    ref = ref_action.build_merge_model(ref_pe, mc['num_actions'], mc['input_shape'], mc['num_frames'], mc['num_joints'],
                                       mc['num_blocks'], pose_dim=2)
    prod_pe = reception.build(mc['input_shape'], **mc['reception'])
    prod = action.build_merge_model(prod_pe, mc['num_actions'], mc['input_shape'], mc['num_frames'], mc['num_joints'],
                                    mc['num_blocks'], pose_dim=2)
    from ref_cases import positive_last_regmap
    pin('merge_model', ref, prod, mc['seed'], x_merge, strip=('td_',),
        weight_hook=lambda t: positive_last_regmap(t, mc['num_blocks']))


if __name__ == '__main__':
    main()
