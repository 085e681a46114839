"""tests/golden/ref_postprocess.npz: outputs of the REFERENCE's own numpy post-processing
(deephar/utils/transform.py::transform_pose_sequence, deephar/measures.py::pckh / mean_distance_error, imported
unmodified from /root/reference on the Keras shim -- deephar/__init__.py imports keras) on seeded inputs.

    python tests/golden/make_postprocess_golden.py
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'keras_shim'))
sys.path.insert(1, os.environ.get('DEEPHAR_REFERENCE', '/root/reference'))
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402

import deephar  # noqa: E402,F401
from deephar.measures import mean_distance_error, pckh  # noqa: E402
from deephar.utils.transform import transform_pose_sequence  # noqa: E402


def main():
    rng = np.random.default_rng(2024)
    n, nj = 37, 16
    # crop affine maps as the data loaders build them: scale + rotation + translation (image -> [0,1] crop)
    ang = rng.uniform(-0.6, 0.6, n)
    sc = rng.uniform(1 / 400.0, 1 / 150.0, n)
    A = np.zeros((n, 3, 3))
    A[:, 0, 0], A[:, 0, 1], A[:, 1, 0], A[:, 1, 1] = sc * np.cos(ang), -sc * np.sin(ang), sc * np.sin(ang), sc * np.cos(ang)
    A[:, 0, 2], A[:, 1, 2], A[:, 2, 2] = rng.uniform(-0.3, 0.3, n), rng.uniform(-0.3, 0.3, n), 1.0
    y_true_crop = rng.uniform(0.1, 0.9, (n, nj, 2))
    y_pred_crop = y_true_crop + rng.normal(0, 0.03, (n, nj, 2))
    head = rng.uniform(20.0, 60.0, (n, 1))
    y_true = transform_pose_sequence(A.copy(), y_true_crop, inverse=True)
    y_true[rng.uniform(size=(n, nj)) < 0.1] = -1e9                    # missing annotations
    y_pred = transform_pose_sequence(A.copy(), y_pred_crop, inverse=True)
    fwd = transform_pose_sequence(A[0].copy(), y_pred_crop, inverse=False)      # single map, forward
    np.savez_compressed(os.path.join(HERE, 'ref_postprocess.npz'), A=A, y_pred_crop=y_pred_crop, head=head, y_true=y_true,
                        y_pred=y_pred, fwd=fwd, pckh05=pckh(y_true, y_pred, head, refp=0.5),
                        pckh02=pckh(y_true, y_pred, head, refp=0.2), mde=mean_distance_error(y_true, y_pred))
    print('wrote ref_postprocess.npz')


if __name__ == '__main__':
    main()
