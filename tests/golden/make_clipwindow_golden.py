"""tests/golden/ref_clip_windows.npz: outputs of the REFERENCE's own clip-sampling and crop-window helpers
(deephar/data/datasets.py::get_clip_frame_index with random_clip=False, deephar/utils/bbox.py::bbox_to_objposwin /
objposwin_to_bbox as data/pennaction.py:118-134 combines them), imported unmodified on the Keras shim.

    python tests/golden/make_clipwindow_golden.py
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
from deephar.data.datasets import get_clip_frame_index  # noqa: E402
from deephar.utils.bbox import bbox_to_objposwin, objposwin_to_bbox  # noqa: E402


def main():
    rng = np.random.default_rng(11)
    cases = [(int(s), int(sub), int(nf)) for s, sub, nf in zip(rng.integers(3, 400, 60), rng.integers(1, 9, 60),
                                                                rng.choice([8, 16], 60))]
    cases += [(16, 6, 16), (15, 1, 16), (8, 2, 8), (1, 4, 16), (200, 6, 16), (97, 4, 8)]
    frames = [get_clip_frame_index(s, sub, nf, random_clip=False) for s, sub, nf in cases]
    # crop windows as data/pennaction.py:118-134 builds them for evaluation (fixed dconf: scale, transx, transy)
    wins = []
    for _ in range(40):
        w, h = int(rng.integers(200, 700)), int(rng.integers(200, 500))
        scale, tx, ty = float(rng.choice([1.0, 0.7, 1.3])), float(rng.integers(-20, 21)), float(rng.integers(-10, 11))
        if rng.uniform() < 0.5:
            bbox = objposwin_to_bbox(np.array([w / 2, h / 2]), (scale * max(w, h), scale * max(w, h)))      # :127-129
            given = np.full(4, np.nan)
        else:
            x0, y0 = rng.uniform(0, w / 2), rng.uniform(0, h / 2)
            given = np.array([x0, y0, x0 + rng.uniform(5, w / 2), y0 + rng.uniform(5, h / 2)])
            bbox = given
        objpos, winsize = bbox_to_objposwin(bbox)                                                            # :131
        if min(winsize) < 32:
            winsize = (32, 32)
        objpos += scale * np.array([tx, ty])                                                                 # :134
        wins.append([w, h, scale, tx, ty] + list(given) + [objpos[0], objpos[1], winsize[0], winsize[1]])
    np.savez_compressed(os.path.join(HERE, 'ref_clip_windows.npz'), cases=np.array(cases, np.int64),
                        frames=np.array([f + [-1] * (16 - len(f)) for f in frames], np.int64), windows=np.array(wins))
    print('wrote ref_clip_windows.npz')


if __name__ == '__main__':
    main()
