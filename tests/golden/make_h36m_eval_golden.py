"""tests/golden/ref_h36m_eval.npz: the per-block mean per-joint position errors (mm) returned by the REFERENCE's own
Human3.6M evaluator (exp/common/h36m_tools.py::eval_human36m_sc_error with deephar/utils/camera.py, imported unmodified
from /root/reference on the Keras shim) for seeded predictions, affine maps, root depths and serialised cameras
(incl. distortion coefficients): inverse crop affine -> absolute depth -> camera inverse projection -> root-centred MPJPE.

    python tests/golden/make_h36m_eval_golden.py
"""
import contextlib
import io
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
import h36m_tools  # noqa: E402
from deephar.utils.camera import Camera  # noqa: E402


class FakeModel(object):
    def __init__(self, preds):
        self.preds = preds
        self.outputs = [None] * len(preds)
        self.input_shape = (None, 256, 256, 3)

    def predict(self, x, batch_size=8, verbose=0):
        return [p.copy() for p in self.preds]


def main():
    rng = np.random.default_rng(36)
    n, nj, nb = 29, 17, 3
    scam, pose_w = [], np.zeros((n, nj, 3))
    uvd = np.zeros((n, nj, 3))
    for i in range(n):
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] = -q[:, 0]
        cam = Camera(q, rng.normal(0, 500, 3), rng.uniform(1100, 1200, 2), rng.uniform(480, 540, 2),
                     rng.normal(0, 1e-3, 2), rng.normal(0, 5e-3, 3))
        scam.append(cam.serialize())
        # a skeleton 4-6 m in front of the camera
        pts_cam = np.concatenate([rng.normal(0, 400, (nj, 2)), rng.uniform(4000, 6000, (nj, 1))], axis=1)
        pose_w[i] = (np.matmul(cam.R_inv, pts_cam.T) + cam.t).T
        uvd[i] = cam.project(pose_w[i])
    scam = np.array(scam)
    # crop affine (image pixels -> [0,1]), root depth, normalised "ground truth" prediction, noisy per-block predictions
    sc = rng.uniform(1 / 700.0, 1 / 400.0, n)
    afmat = np.zeros((n, 3, 3))
    afmat[:, 0, 0] = afmat[:, 1, 1] = sc
    afmat[:, 0, 2], afmat[:, 1, 2], afmat[:, 2, 2] = rng.uniform(-0.4, 0.1, n), rng.uniform(-0.4, 0.1, n), 1.0
    rootz = uvd[:, 0, 2].copy()
    resol_z = 2000.0
    ideal = np.zeros((n, nj, 3))
    ideal[:, :, 0:2] = np.einsum('nij,nkj->nki', afmat[:, :2, :2], uvd[:, :, 0:2]) + afmat[:, None, :2, 2]
    ideal[:, :, 2] = (uvd[:, :, 2] - rootz[:, None]) / resol_z + 0.5
    preds = [np.concatenate([ideal + rng.normal(0, 0.02 / (b + 1), ideal.shape), rng.uniform(size=(n, nj, 1))], axis=-1)
             .astype(np.float32) for b in range(nb)]
    action = rng.integers(0, 15, (n, 1))
    # the dataset loader sets this module global when annotations are read (data/human36m.py:58-59)
    import deephar.data.human36m as h36m_data
    h36m_data.ACTION_LABELS = ['action%02d' % i for i in range(15)]
    x = np.zeros((n, 1, 1, 3), np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        scores = h36m_tools.eval_human36m_sc_error(FakeModel(preds), x, pose_w.copy(), afmat.copy(), rootz.copy(), scam,
                                                   action, resol_z=resol_z, batch_size=8, verbose=False)
    np.savez_compressed(os.path.join(HERE, 'ref_h36m_eval.npz'), scores=np.asarray(scores, np.float64), preds=np.stack(preds),
                        afmat=afmat, rootz=rootz, scam=scam, pose_w=pose_w, resol_z=resol_z)
    print('wrote ref_h36m_eval.npz; scores (mm)', scores)


if __name__ == '__main__':
    main()
