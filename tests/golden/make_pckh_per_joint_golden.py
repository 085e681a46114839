"""tests/golden/ref_pckh_per_joint.npz: the per-joint PCKh table the REFERENCE prints at the end of its MPII evaluation
(deephar/measures.py::pckh_per_joint, imported unmodified from /root/reference on the Keras shim; it returns nothing, so
the printed ' %.2f | ' cells are parsed) for the seeded poses of ref_postprocess.npz, at refp 0.5 and 0.2.

    python tests/golden/make_pckh_per_joint_golden.py
"""
import contextlib
import io
import os
import re
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'keras_shim'))
sys.path.insert(1, os.environ.get('DEEPHAR_REFERENCE', '/root/reference'))
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402

import deephar  # noqa: E402,F401
from deephar.measures import pckh_per_joint  # noqa: E402
from deephar.utils.pose import pa16j2d  # noqa: E402


def main():
    z = np.load(os.path.join(HERE, 'ref_postprocess.npz'))
    out = {}
    for refp in (0.5, 0.2):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            pckh_per_joint(z['y_true'], z['y_pred'], z['head'], pa16j2d, refp=refp, verbose=1)
        text = re.sub(r'\x1b\[[0-9;]*m', '', buf.getvalue())
        cells = [float(v) for v in re.findall(r'(-?\d+\.\d\d) \|', text.splitlines()[-1])]
        assert len(cells) == 16, text
        out['percent_%s' % refp] = np.array(cells)
    out['joint_names'] = np.array(pa16j2d.joint_names)
    np.savez_compressed(os.path.join(HERE, 'ref_pckh_per_joint.npz'), **out)
    print('wrote ref_pckh_per_joint.npz', out['percent_0.5'])


if __name__ == '__main__':
    main()
