"""tests/golden/ref_bbox_from_poses.npz: bounding boxes returned by the REFERENCE's own
exp/common/generic.py::get_bbox_from_poses (with deephar/utils/bbox.py::get_valid_bbox_array and
deephar/utils/transform.py::transform_2d_points, imported unmodified from /root/reference on the Keras shim) for seeded
frame- and clip-shaped predictions: joints kept where the confidence column exceeds its threshold, 1.5x square box
around them per frame, union over the frames, mapped back through the inverse crop affine.

    python tests/golden/make_bbox_golden.py
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('DEEPHAR_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.join(HERE, 'keras_shim'))
sys.path.insert(1, REF)
sys.path.insert(2, os.path.join(REF, 'exp', 'common'))
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402

import deephar  # noqa: E402,F401
import generic  # noqa: E402


def main():
    rng = np.random.default_rng(2018)
    cases = {}
    for name, shape in (('frames', (5, 16, 3)), ('clip', (1, 8, 16, 3)), ('frames3d', (3, 17, 4))):
        poses = rng.uniform(0.1, 0.9, shape)
        poses[..., -1] = rng.uniform(0.0, 1.0, shape[:-1])            # visibility column
        ang = rng.uniform(-0.3, 0.3)
        sc = rng.uniform(1 / 400.0, 1 / 150.0)
        afmat = np.array([[sc * np.cos(ang), -sc * np.sin(ang), rng.uniform(-0.4, 0.1)],
                          [sc * np.sin(ang), sc * np.cos(ang), rng.uniform(-0.4, 0.1)], [0, 0, 1.0]])
        for scale in (1.5, 1.2):
            cases['%s_%s' % (name, scale)] = (poses, afmat, scale, generic.get_bbox_from_poses(poses, afmat, scale=scale))
    out = {}
    for k, (poses, afmat, scale, bbox) in cases.items():
        out[k + '/poses'], out[k + '/afmat'], out[k + '/scale'], out[k + '/bbox'] = poses, afmat, scale, bbox
    np.savez_compressed(os.path.join(HERE, 'ref_bbox_from_poses.npz'), **out)
    print('wrote ref_bbox_from_poses.npz:', {k: np.round(v[3], 2).tolist() for k, v in cases.items()})


if __name__ == '__main__':
    main()
