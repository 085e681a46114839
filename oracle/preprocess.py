"""CPU restatement of the evaluation-time input pipeline (TEST INFRASTRUCTURE, see oracle/__init__.py):

  deephar/utils/transform.py:60-134  T.rotate_crop(angle = 0) -> T.crop(integer box, zero fill outside the image)
                                     -> T.resize(crop_resolution, Image.BILINEAR) [-> T.horizontal_flip()] -> T.asarray()
  deephar/utils/transform.py:212-231 normalize_channels: x / 255, ** chpower, (x - 0.5) * 2        (float32)
  as driven by deephar/data/mpii.py:91-122 with the fixed (evaluation) data configuration.

The arithmetic of `Image.resize(size, BILINEAR)` lives in Pillow (not vendored; any version >= 7: `ImagingResample`,
libImaging/Resample.c), restated here from its published algorithm and PINNED bit-for-bit against the Pillow installed
in this image by tests/test_preprocess.py:
  per axis  scale = in / out, filterscale = max(scale, 1), support = 1.0 * filterscale   (triangle filter)
            for every output index xx: center = (xx + 0.5) * scale, window [xmin, xmin + n) =
            [int(center - support + 0.5), int(center + support + 0.5)) clipped to the image, weights
            w_i = tri((i + xmin - center + 0.5) / filterscale) normalised to sum 1, then fixed-point
            k_i = int(0.5 + w_i * 2^22)                                   (8-bit images: PRECISION_BITS = 22)
  pixels    horizontal pass then vertical pass, each: clip8((2^21 + sum_i pixel_i * k_i) >> 22), uint8 in between.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resample_coefficients(in_size, out_size):
    """-> (bounds int32 (out, 2): first source index and tap count; coefs int32 (out, ksize)) -- Resample.c
    precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter on the full-image box."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale > 1.0 else 1.0
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    coefs = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = np.zeros(ksize, np.float64)
        ww = 0.0
        for x in range(n):
            v = (x + xmin - center + 0.5) * ss
            if v < 0.0:
                v = -v
            wt = 1.0 - v if v < 1.0 else 0.0
            w[x] = wt
            ww += wt
        if ww != 0.0:
            for x in range(n):
                w[x] /= ww
        for x in range(ksize):
            coefs[xx, x] = int(0.5 + w[x] * (1 << PRECISION_BITS)) if w[x] >= 0 else int(-0.5 + w[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, n)
    return bounds, coefs


def _pass(img, bounds, coefs, axis):
    """One resampling pass over `axis` (0 = rows / vertical, 1 = columns / horizontal) of a uint8 (H, W, C) image."""
    src = np.moveaxis(img.astype(np.int64), axis, 0)
    out = np.empty((len(bounds),) + src.shape[1:], np.int64)
    for i, (first, n) in enumerate(bounds):
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(n):
            acc += src[first + t] * int(coefs[i, t])
        out[i] = acc >> PRECISION_BITS
    return np.moveaxis(np.clip(out, 0, 255).astype(np.uint8), 0, axis)


def crop(img, box):
    """PIL Image.crop((x0, y0, x1, y1)): pixels outside the image are 0."""
    x0, y0, x1, y1 = [int(v) for v in box]
    h, w = img.shape[:2]
    out = np.zeros((max(y1 - y0, 0), max(x1 - x0, 0), img.shape[2]), img.dtype)
    sx0, sy0, sx1, sy1 = max(x0, 0), max(y0, 0), min(x1, w), min(y1, h)
    if sx1 > sx0 and sy1 > sy0:
        out[sy0 - y0:sy1 - y0, sx0 - x0:sx1 - x0] = img[sy0:sy1, sx0:sx1]
    return out


def resize_bilinear(img, size):
    """PIL Image.resize((w, h), Image.BILINEAR) of a uint8 (H, W, C) image: horizontal pass, then vertical."""
    w_out, h_out = size
    if img.shape[1] != w_out:
        img = _pass(img, *resample_coefficients(img.shape[1], w_out), axis=1)
    if img.shape[0] != h_out:
        img = _pass(img, *resample_coefficients(img.shape[0], h_out), axis=0)
    return img


def normalize_channels(frame, channel_power=1):
    """transform.py:212-231 on a float32 array (in place semantics of the reference: float32 throughout)."""
    frame = np.array(frame, dtype=np.float32)
    frame /= np.float32(255.)
    if isinstance(channel_power, int):
        if channel_power != 1:
            frame = np.power(frame, channel_power)
    else:
        for c in range(3):
            if channel_power[c] != 1:
                frame[:, :, c] = np.power(frame[:, :, c], channel_power[c])
    frame -= np.float32(.5)
    frame *= np.float32(2.)
    return frame


def eval_frame(img, box, size=(256, 256), hflip=False, channel_power=1):
    """mpii.py:107-122 with angle = 0: crop -> resize -> [flip] -> normalize.  img: uint8 (H, W, 3)."""
    out = resize_bilinear(crop(img, box), size)
    if hflip:
        out = out[:, ::-1]
    return normalize_channels(out, channel_power)
