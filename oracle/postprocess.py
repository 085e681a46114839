"""CPU restatement of the evaluator-side post-processing (TEST INFRASTRUCTURE, see oracle/__init__.py):
deephar/utils/transform.py:136-209 (transform_2d_points, transform_pose_sequence) and
deephar/measures.py:5-76 (_valid_joints, mean_distance_error, pckh), numpy fp64."""
import numpy as np

PCKH_USED_JOINTS = [2, 3, 4, 5, 6, 7, 10, 11, 12, 13, 14, 15, 8, 9]        # measures.py:65


def transform_pose_sequence(A, poses, inverse=True):
    """transform.py:174-209."""
    A = np.array(A, dtype=np.float64)
    poses = np.asarray(poses, dtype=np.float64)
    assert poses.ndim == 3
    if inverse:
        A = np.linalg.inv(A)                       # batched for (N,3,3), as the reference's per-sample loop
    y = np.empty(poses.shape[:2] + (2,))
    for j in range(len(poses)):
        M = A[j] if A.ndim == 3 else A
        h = np.ones((3, poses.shape[1]))
        h[0:2, :] = poses[j, :, 0:2].T
        y[j] = np.dot(M, h)[0:2].T                 # transform_2d_points (transform.py:136-171)
    return y


def _valid(y, min_valid=-1e6):
    return np.all(y > min_valid, axis=-1).astype(np.float64)


def pckh(y_true, y_pred, head_size, refp=0.5):
    """measures.py:49-76."""
    y_true = np.asarray(y_true, np.float64)[:, PCKH_USED_JOINTS, :]
    y_pred = np.asarray(y_pred, np.float64)[:, PCKH_USED_JOINTS, :]
    valid = _valid(y_true)
    dist = np.sqrt(np.sum((y_true - y_pred) ** 2, axis=-1)) / np.asarray(head_size, np.float64).reshape(-1, 1)
    return float(((dist <= refp) * valid).sum() / valid.sum())


def pckh_per_joint(y_true, y_pred, head_size, refp=0.5):
    """measures.py:108-146 without the printing: per joint, the share of valid samples within refp x head size."""
    y_true, y_pred = np.asarray(y_true, np.float64), np.asarray(y_pred, np.float64)
    valid = np.stack([_valid(y) for y in y_true]).astype(np.float64)
    dist = np.sqrt(((y_true - y_pred) ** 2).sum(axis=-1)) / np.asarray(head_size, np.float64).reshape(-1, 1)
    with np.errstate(invalid='ignore', divide='ignore'):
        return ((dist <= refp) * valid).sum(axis=0) / valid.sum(axis=0)


def mean_distance_error(y_true, y_pred):
    """measures.py:18-47."""
    y_true, y_pred = np.asarray(y_true, np.float64), np.asarray(y_pred, np.float64)
    valid = _valid(y_true)
    dist = np.sqrt(np.sum((y_true - y_pred) ** 2, axis=-1))
    return float((dist * valid).sum() / valid.sum())


def multiclip_action_scores(probs, video_of_item, truth, n_videos=None):
    """exp/common/penn_tools.py:85-150 (eval_multiclip_dataset) after the predictions: per video the PRODUCT over its
    clips x {no flip, h-flip} of every block's action probabilities (float64 accumulator starting at 1), arg-max against
    the label, accuracy in percent per block.  probs (num_blocks, N_items, n_act); plain loops as in the reference."""
    probs = np.asarray(probs)
    nb, n_items, n_act = probs.shape
    n_videos = int(n_videos if n_videos is not None else np.max(video_of_item) + 1)
    a_pred = np.ones((nb, n_videos, n_act))
    for k in range(n_items):
        for b in range(nb):
            a_pred[b, video_of_item[k], :] *= probs[b][k]
    correct = np.argmax(a_pred, axis=-1) == np.asarray(truth)[None, :]
    return 100 * np.sum(correct, axis=-1) / n_videos


def camera_inverse_project(scam, uvd):
    """deephar/utils/camera.py:52-71 + camera_deserialize (:97-110) for ONE serialised camera
    [R(9) t(3) f(2) c(2) p(2) k(3, optional)]: (u, v pixels, depth mm) -> world mm.  Like the reference, the working
    copy keeps the dtype of `uvd` (float32 predictions stay float32 through the in-place steps)."""
    scam = np.asarray(scam, np.float64)
    R, t = scam[0:9].reshape(3, 3), scam[9:12].reshape(3, 1)
    f, c, p = scam[12:14].reshape(1, 2), scam[14:16].reshape(1, 2), scam[16:18].reshape(1, 2)
    k = scam[18:21] if len(scam) > 18 else None
    x = uvd.copy()
    x[:, 0:2] = (x[:, 0:2] - c) / f
    if k is not None:
        r2 = np.power(x[:, 0], 2) + np.power(x[:, 1], 2)
        radial = 1. + r2 * k[0] + np.power(r2, 2) * k[1] + np.power(r2, 3) * k[2]
        tan = np.sum(x[:, 0:2] * p, axis=-1)
        x[:, 0:2] -= np.dot(np.expand_dims(r2, axis=-1), p)
        x[:, 0:2] /= np.expand_dims(radial + tan, axis=-1)
    x[:, 0:2] *= x[:, 2:3]
    return (np.matmul(np.linalg.inv(R), x.T) + t).T


def human36m_mpjpe(preds, afmat, rootz, scam, pose_w, resol_z=2000.):
    """exp/common/h36m_tools.py:12-99 after the predict: per block, (x, y) back through the inverse crop affine,
    z = resol_z * (z - 0.5) + rootz, camera inverse projection per sample, both poses root-centred, mean joint
    distance in mm.  preds: list of (N, nj, >= 3); per-sample loop and in-place (dtype-preserving) updates as in the
    reference, so float32 predictions are rounded to float32 after every step exactly as there."""
    y_true = np.array(pose_w, np.float64)
    y_true -= y_true[:, 0:1, :]
    rootz = np.asarray(rootz, np.float64).reshape(-1, 1)
    out = []
    for p in preds:
        y = np.array(p)[:, :, 0:3]
        y[:, :, 0:2] = transform_pose_sequence(np.array(afmat, np.float64), y[:, :, 0:2], inverse=True)
        y[:, :, 2] = (resol_z * (y[:, :, 2] - 0.5)) + rootz
        w = np.zeros(y_true.shape)
        for j in range(len(y)):
            w[j] = camera_inverse_project(scam[j], y[j])
        w -= w[:, 0:1, :]
        out.append(mean_distance_error(y_true, w))
    return out
