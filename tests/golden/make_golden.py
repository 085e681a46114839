"""Generates the committed fixtures under tests/golden/.  Run from the repo root IN THE
AUTHORING CONTAINER (it reads /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

1. linspace_2d.npz -- produced by EXECUTING the reference's own source text of
   deephar/utils/math.py::linspace_2d (lines 6-19, pure numpy; the module itself cannot be
   imported because it imports keras).  This is the only piece of the reference that can
   run here; it pins the soft-argmax coordinate grid.
2. reception_small_oracle.npz -- outputs of the fp64 oracle (oracle/reception.py) on a small
   seeded configuration; a regression pin for the oracle itself (NOT reference output).
"""
import ast
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF_MATH = '/root/reference/deephar/utils/math.py'


def reference_linspace_2d():
    src = open(REF_MATH).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'linspace_2d'][0]
    code = ast.get_source_segment(src, fn)
    ns = {'np': np}
    exec(compile(code, REF_MATH, 'exec'), ns)
    return ns['linspace_2d']


def main():
    f = reference_linspace_2d()
    out = {}
    for (r, c) in [(4, 5), (32, 32), (16, 16), (8, 8), (4, 4), (3, 7)]:
        for dim in (0, 1):
            out['r%d_c%d_d%d' % (r, c, dim)] = f(r, c, dim=dim)
    np.savez(os.path.join(HERE, 'linspace_2d.npz'), **out)
    print('linspace_2d.npz', len(out), 'arrays')

    from oracle import ops_np, reception, synth
    from deephar_b200.weights import load_calibration
    tab = synth.SyntheticTable(1234, load_calibration('reception_j16_d2_c2_k5'))
    x = synth.synth_frames(1, 64, 64, seed=3)
    outs = reception.forward(ops_np, tab, x, 16, 2, num_context_per_joint=2, num_blocks=2, ksize=(5, 5),
                             concat_pose_confidence=True)
    np.savez(os.path.join(HERE, 'reception_small_oracle.npz'), *[o.astype(np.float64) for o in outs])
    print('reception_small_oracle.npz', [o.shape for o in outs])


if __name__ == '__main__':
    main()
