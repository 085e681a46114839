"""tests/golden/ref_action_eval.npz: scores of the REFERENCE's own multi-clip action evaluator
(exp/common/penn_tools.py::eval_multiclip_dataset, imported unmodified from /root/reference on the Keras shim) run on a
stand-in dataset / model that return seeded probabilities: the per-video product over clips x {no flip, h-flip} of
every prediction block's action probabilities, arg-max, accuracy.

    python tests/golden/make_action_eval_golden.py
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
import penn_tools  # noqa: E402


class _Conf(object):
    fixed_hflip = 0


class FakePenn(object):
    """The slice of the PennAction loader the evaluator touches."""

    def __init__(self, clips_per_video, n_act, truth):
        self.dataconf = _Conf()
        self.clips, self.n_act, self.truth = clips_per_video, n_act, truth

    def get_length(self, mode):
        return len(self.clips)

    def get_shape(self, key):
        assert key == 'pennaction'
        return (self.n_act,)

    def get_clip_index(self, i, mode, subsamples):
        return [[c] for c in range(self.clips[i])]

    def get_data(self, i, mode, frame_list):
        onehot = np.zeros(self.n_act)
        onehot[self.truth[i]] = 1
        # the "frame" carries the identity of the item so that the fake model can look its prediction up
        return {'pennaction': onehot, 'frame': np.array([i, frame_list[0], self.dataconf.fixed_hflip], np.float64)}


class FakeModel(object):
    def __init__(self, table, num_blocks):
        self.table, self.outputs = table, [None] * num_blocks

    def predict(self, x):
        i, c, h = [int(v) for v in x[0]]
        return [self.table[b][(i, c, h)][None] for b in range(len(self.outputs))]


def main():
    rng = np.random.default_rng(77)
    n_videos, n_act, num_blocks = 23, 15, 4
    clips = rng.integers(1, 6, n_videos)
    truth = rng.integers(0, n_act, n_videos)
    table = [dict() for _ in range(num_blocks)]
    items, probs = [], [[] for _ in range(num_blocks)]
    for i in range(n_videos):
        for c in range(clips[i]):
            for h in range(2):
                items.append((i, c, h))
                for b in range(num_blocks):
                    logit = rng.normal(0, 1.0, n_act)
                    logit[truth[i]] += 0.4 * (b + 1)                     # later blocks are better
                    p = np.exp(logit) / np.exp(logit).sum()
                    p = p.astype(np.float32)
                    table[b][(i, c, h)] = p
                    probs[b].append(p)
    penn = FakePenn(clips, n_act, truth)
    scores = penn_tools.eval_multiclip_dataset(FakeModel(table, num_blocks), penn, subsampling=1, verbose=0)
    np.savez_compressed(os.path.join(HERE, 'ref_action_eval.npz'), scores=np.asarray(scores, np.float64),
                        video_of_item=np.array([it[0] for it in items], np.int64),
                        probs=np.stack([np.stack(p) for p in probs]), truth=truth.astype(np.int64))
    print('wrote ref_action_eval.npz; scores', scores)


if __name__ == '__main__':
    main()
