"""tests/golden/ref_preprocess.npz: frames produced by the REFERENCE's own input pipeline
(deephar/utils/transform.py::T.rotate_crop / resize / horizontal_flip / normalize_affinemap + normalize_channels, driven
as deephar/data/mpii.py:91-122 does with the fixed evaluation config: angle 0) on small synthetic images.
Runs where /root/reference and Pillow exist; deephar/__init__.py imports keras -> the Keras shim.

    python tests/golden/make_preprocess_golden.py
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'keras_shim'))
sys.path.insert(1, os.environ.get('DEEPHAR_REFERENCE', '/root/reference'))
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
from PIL import Image  # noqa: E402

import deephar  # noqa: E402,F401
from deephar.utils.transform import T, normalize_channels  # noqa: E402


def synth_image(h, w, seed):
    """smooth, compressible RGB test image with edges (gradients + rectangles + a little noise)"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // max(h + w - 2, 1))], -1)
    for _ in range(6):
        x0, y0 = rng.integers(0, w - 8), rng.integers(0, h - 8)
        x1, y1 = x0 + rng.integers(4, w // 3), y0 + rng.integers(4, h // 3)
        img[y0:y1, x0:x1] = rng.integers(0, 256, 3)
    img = img + rng.integers(-6, 7, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


CASES = [
    # name, (h, w), objpos (x, y), winsize, crop_resolution, hflip
    ('down', (180, 240), (118.3, 96.7), 150.6, (64, 64), 0),
    ('down_flip', (180, 240), (60.2, 40.9), 140.0, (64, 64), 1),          # window leaves the image: zero fill
    ('up', (90, 120), (61.5, 44.2), 40.4, (96, 96), 0),                   # upscaling
    ('rect', (200, 150), (70.0, 110.0), 180.9, (48, 80), 1),              # non-square resolution (w, h) = (48, 80)
]


def main():
    blob = {}
    for name, (h, w), objpos, win, res, hflip in CASES:
        src = synth_image(h, w, seed=len(name) * 7 + h)
        imgt = T(Image.fromarray(src))
        imgt.rotate_crop(0, np.array(objpos), (win, win))           # mpii.py:114
        imgt.resize(res)                                            # mpii.py:115
        if hflip == 1:
            imgt.horizontal_flip()                                  # mpii.py:117-118
        imgt.normalize_affinemap()
        frame = normalize_channels(imgt.asarray(), channel_power=1)   # mpii.py:121-122
        blob[name + '_src'] = src
        blob[name + '_args'] = np.array([objpos[0], objpos[1], win, res[0], res[1], hflip], np.float64)
        blob[name + '_frame'] = frame.astype(np.float32)
        blob[name + '_afmat'] = imgt.afmat.copy()
    np.savez_compressed(os.path.join(HERE, 'ref_preprocess.npz'), **blob)
    print('wrote ref_preprocess.npz (%d KB)' % (os.path.getsize(os.path.join(HERE, 'ref_preprocess.npz')) // 1024))


if __name__ == '__main__':
    main()
