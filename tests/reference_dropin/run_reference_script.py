"""Runs one of the reference's ENTRY SCRIPTS, unmodified (runpy), on deephar_b200 after dropin.install():

    mpii : exp/mpii/eval_mpii_singleperson.py  (BASELINE configs[0]/[1]: the headline model's evaluator)
    h36m : exp/h36m/eval_h36m.py               (BASELINE configs[2])

Everything model-side of the script is the product's: `reception.build` records the deephar_b200 model,
`get_file` finds the weight file in the Keras cache, `model.load_weights` reads the Keras HDF5 file, the script's
`Model(model.input, [concatenate([pose_b, vis_b]) ...])` re-wrap (made AFTER load_weights) compiles a second plan on
the same layers and keeps their loaded weights, and the reference's own evaluator (exp/common/*_tools.py) drives
`model.predict`.  What is NOT the product's, because it does not exist in the container: the dataset (a seeded
stand-in for `deephar.data.*` returning the arrays the script unpacks), the released checkpoint (seeded weights in the
same Keras HDF5 layout, written by deephar_b200.keras_h5.save) and -- on CPU only -- the forward itself: without a
GPU `Model.predict` is the oracle's forward behind the product's own input checks (TEST INFRASTRUCTURE; set
DEEPHAR_B200_SCRIPT_ON_GPU=1 on a B200 to run the real one).  Prints one JSON line: the evaluator's scores, the same
scores recomputed from the oracle's outputs with the oracle's post-processing, and what the compiled model looks like.
"""
import contextlib
import io
import json
import os
import runpy
import sys
import types
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('DEEPHAR_REFERENCE', '/root/reference')
sys.path[:0] = [REF, ROOT]
warnings.filterwarnings('ignore')
import numpy as np  # noqa: E402

import deephar_b200.dropin  # noqa: E402
deephar_b200.dropin.install()
_stderr, sys.stderr = sys.stderr, open(os.devnull, 'w')       # the reference prints a banner on import
import deephar  # noqa: E402,F401
import deephar.data  # noqa: E402
sys.stderr = _stderr

from deephar_b200 import keras_h5  # noqa: E402
from deephar_b200 import model as product_model  # noqa: E402
from deephar_b200 import reception as product_reception  # noqa: E402
from oracle import ops_torch, synth  # noqa: E402
from oracle import postprocess as oracle_post  # noqa: E402
from oracle import reception as oracle_reception  # noqa: E402

ON_GPU = os.environ.get('DEEPHAR_B200_SCRIPT_ON_GPU') == '1'
N = 3

CASES = {
    'mpii': dict(script='exp/mpii/eval_mpii_singleperson.py', weights='weights_PE_MPII_cvpr18_19-09-2017.h5',
                 kw=dict(num_joints=16, dim=2, num_blocks=8, num_context_per_joint=2, ksize=(5, 5),
                         concat_pose_confidence=False)),
    'h36m': dict(script='exp/h36m/eval_h36m.py', weights='weights_3DPE_H36M_cvpr18_Nov-2017.h5',
                 kw=dict(num_joints=17, dim=3, num_blocks=8, ksize=(5, 5), concat_pose_confidence=False)),
}


def oracle_outputs(kw, table, x):
    return [np.asarray(o, np.float32) for o in oracle_reception.forward(ops_torch, table, x, **kw)]


def install_oracle_predict(kw):
    """CPU stand-in for the device forward: the product's own input handling, then the oracle (see the module docstring)."""
    def predict(self, x, batch_size=32, verbose=0):
        x = self._host_input(x)
        outs = oracle_outputs(kw, self.get_weights(), x)
        if len(self.graph.outputs) == len(outs) // 2:           # the script's re-wrapped model: [pose_b | vis_b] per block
            outs = [np.concatenate([outs[2 * b], outs[2 * b + 1]], axis=-1) for b in range(len(outs) // 2)]
        assert [tuple(o.shape[1:]) for o in outs] == [tuple(s[1:]) for s in self.output_shape]
        return outs[0] if len(outs) == 1 else outs
    product_model.Model.predict = predict


class FakeDataset(object):
    def __init__(self, *args, **kwargs):
        pass

    def get_length(self, mode):
        return N


def mpii_batches():
    rng = np.random.default_rng(11)
    x = synth.synth_frames(N, 256, 256, seed=21)
    pose = np.concatenate([rng.uniform(0.2, 0.8, (N, 16, 2)), np.ones((N, 16, 1))], axis=-1)
    afmat = np.tile(np.array([[1 / 300.0, 0, 0.05], [0, 1 / 300.0, 0.1], [0, 0, 1]]), (N, 1, 1))
    head = rng.uniform(40, 60, (N, 1))
    return [x], [pose, afmat, head]


def h36m_batches():
    from deephar.utils.camera import Camera
    rng = np.random.default_rng(36)
    x = synth.synth_frames(N, 256, 256, seed=22)
    nj = 17
    scam, pose_w, uvd = [], np.zeros((N, nj, 3)), np.zeros((N, nj, 3))
    for i in range(N):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        cam = Camera(q, rng.normal(0, 500, 3), rng.uniform(1100, 1200, 2), rng.uniform(480, 540, 2),
                     rng.normal(0, 1e-3, 2), rng.normal(0, 5e-3, 3))
        scam.append(cam.serialize())
        pts = np.concatenate([rng.normal(0, 400, (nj, 2)), rng.uniform(4000, 6000, (nj, 1))], axis=1)
        pose_w[i] = (np.matmul(cam.R_inv, pts.T) + cam.t).T
        uvd[i] = cam.project(pose_w[i])
    afmat = np.zeros((N, 3, 3))
    afmat[:, 0, 0] = afmat[:, 1, 1] = 1 / 500.0
    afmat[:, 0, 2], afmat[:, 1, 2], afmat[:, 2, 2] = -0.3, -0.2, 1.0
    import deephar.data.human36m as h36m_data
    h36m_data.ACTION_LABELS = ['action%02d' % i for i in range(15)]      # set by the loader when it reads the annotations
    return [x], [pose_w, uvd, afmat, np.array(scam), rng.integers(0, 15, (N, 1))]


def main():
    case = CASES[sys.argv[1]]
    kw = case['kw']
    # the "released checkpoint": seeded weights in the Keras HDF5 layout, in the Keras cache where get_file looks
    cache = os.path.join(os.path.expanduser(os.environ.get('KERAS_HOME') or os.path.join('~', '.keras')), 'models')
    os.makedirs(cache, exist_ok=True)
    src = product_reception.build((256, 256, 3), **kw).init_synthetic_weights(77)
    table = src.get_weights()
    keras_h5.save(os.path.join(cache, case['weights']), src.weight_specs, table)

    batches = mpii_batches() if sys.argv[1] == 'mpii' else h36m_batches()

    class FakeLoader(object):
        def __init__(self, *args, **kwargs):
            pass

        def __getitem__(self, i):
            return batches

    for name in ('MpiiSinglePerson', 'Human36M'):
        setattr(deephar.data, name, FakeDataset)
    deephar.data.BatchLoader = FakeLoader
    sys.modules['annothelper'] = types.ModuleType('annothelper')
    for fn in ('check_mpii_dataset', 'check_h36m_dataset', 'check_pennaction_dataset', 'check_ntu_dataset'):
        setattr(sys.modules['annothelper'], fn, lambda: None)
    np.float = float                                            # the reference predates numpy 1.24
    if not ON_GPU:
        install_oracle_predict(kw)

    os.chdir(REF)
    sys.argv = [os.path.join(REF, case['script'])]
    printed = io.StringIO()
    with contextlib.redirect_stdout(printed):
        g = runpy.run_path(sys.argv[0], run_name='__main__')

    model = g['model']
    impl = model._compiled()
    held = impl.get_weights()
    x = batches[0][0]
    # the same numbers from the oracle alone: forward on the file's weights + the oracle's post-processing
    ref = oracle_outputs(kw, table, x)
    with contextlib.redirect_stdout(io.StringIO()):
        if sys.argv[0].endswith('eval_mpii_singleperson.py'):
            scores = g['eval_singleperson_pckh'](model, g['x_val'], g['p_val'][:, :, 0:2], g['afmat_val'], g['head_val'],
                                                 verbose=0)
            y_true = oracle_post.transform_pose_sequence(g['afmat_val'], g['p_val'][:, :, 0:2], inverse=True)
            want = [oracle_post.pckh(y_true, oracle_post.transform_pose_sequence(g['afmat_val'], ref[2 * b][:, :, 0:2],
                                                                                 inverse=True), g['head_val'], 0.5)
                    for b in range(len(ref) // 2)]
        else:
            scores = g['eval_human36m_sc_error'](model, g['x_val'], g['pw_val'], g['afmat_val'], g['puvd_val'][:, 0, 2],
                                                 g['scam_val'], g['action'], batch_size=24, verbose=False)
            want = oracle_post.human36m_mpjpe([ref[2 * b] for b in range(len(ref) // 2)], g['afmat_val'],
                                              g['puvd_val'][:, 0, 2], g['scam_val'], g['pw_val'])
    print(json.dumps({
        'scores': [float(s) for s in scores], 'oracle_scores': [float(s) for s in want],
        'model_class': type(model).__module__ + '.' + type(model).__name__, 'n_outputs': len(model.outputs),
        'output_shape': [list(s) for s in model.output_shape],
        'weights_are_the_files': all(np.array_equal(held[k], table[k]) for k in table) and set(held) == set(table),
        'launches': len(impl.plan.kops), 'kinds_tail': [k.kind for k in impl.plan.kops][-2:],
        'script_printed': printed.getvalue()[-400:], 'forward': 'B200' if ON_GPU else 'oracle (CPU stand-in)'}))


if __name__ == '__main__':
    main()
