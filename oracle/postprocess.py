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
