"""Weight containers for the deephar_b200 models.

Layouts are the reference's Keras layouts, unchanged: Conv2D kernel
(kh,kw,Cin,Cout); SeparableConv2D depthwise_kernel (kh,kw,Cin,1) + pointwise_kernel
(1,1,Cin,Cout); BatchNormalization [gamma] beta moving_mean moving_variance; no conv
biases (deephar/layers.py:69,78).  Names are "<sub-model>/<layer>/<weight>".

Also holds the seeded synthetic-weight recipe used by bench.py and the tests (there is
no network to fetch the released .h5 files): SURVEY.md 8(d), plus a small per-layer
calibration table (deephar_b200/synth_calib/*.json) that makes the synthetic BatchNorm
statistics normalise like trained ones (see tests/golden/make_calibration.py).
"""
import json
import os
import zlib

import numpy as np

BN_EPS = 1e-3          # keras BatchNormalization default epsilon
_CALIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'synth_calib')


def load_calibration(key):
    """Per-layer calibration scalars for a model family (empty dict if none committed)."""
    path = os.path.join(_CALIB_DIR, key + '.json')
    if not os.path.exists(path):
        return {}
    with open(path) as f:
        raw = json.load(f)
    return {k: (tuple(v) if isinstance(v, list) else v) for k, v in raw.items()}


def synthetic_weight(seed, name, shape, calib=None):
    """Deterministic synthetic value for weight `name` (depends only on seed/name/shape/calib)."""
    calib = calib or {}
    rng = np.random.default_rng([int(seed), zlib.crc32(name.encode('utf-8'))])
    layer, leaf = name.rsplit('/', 1)
    if leaf == 'kernel':
        std = np.sqrt(2.0 / (shape[0] * shape[1] * shape[2])) * calib.get(name, 1.0)
        a = rng.standard_normal(shape) * std
    elif leaf == 'pointwise_kernel':
        a = rng.standard_normal(shape) * np.sqrt(2.0 / shape[2])
    elif leaf == 'depthwise_kernel':
        a = rng.standard_normal(shape) * np.sqrt(1.0 / (shape[0] * shape[1]))
    elif leaf == 'beta':
        a = rng.standard_normal(shape) * 0.1
    elif leaf == 'moving_mean':
        m, v = calib.get(layer, (0.0, 1.0))
        a = m + rng.standard_normal(shape) * 0.1 * np.sqrt(v)
    elif leaf == 'moving_variance':
        m, v = calib.get(layer, (0.0, 1.0))
        a = v * rng.uniform(0.5, 1.5, shape)
    elif leaf == 'gamma':
        a = rng.uniform(0.8, 1.2, shape)
    else:
        raise KeyError('unknown weight kind: %s' % name)
    return a.astype(np.float32)


def synthetic_weights(specs, seed=1234, calib=None):
    """specs: ordered [(name, shape)] -> {name: float32 array}."""
    return {name: synthetic_weight(seed, name, tuple(shape), calib) for name, shape in specs}


def fold_batchnorm(gamma, beta, mean, var):
    """y = (x-mean)/sqrt(var+eps)*gamma + beta  ->  y = x*scale + shift (fp64 fold, fp32 result)."""
    inv = 1.0 / np.sqrt(var.astype(np.float64) + BN_EPS)
    scale = inv if gamma is None else inv * gamma.astype(np.float64)
    shift = beta.astype(np.float64) - mean.astype(np.float64) * scale
    return scale.astype(np.float32), shift.astype(np.float32)


def split_bf16(w):
    """fp32 -> (hi, lo) bf16 pair with hi + lo ~= w to ~16 mantissa bits (round-to-nearest-even).
    Returned as uint16 bit patterns."""
    w = np.ascontiguousarray(w, dtype=np.float32)

    def to_bf16_bits(a):
        u = a.view(np.uint32).astype(np.uint64)
        rounded = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
        return rounded.astype(np.uint16)

    hi = to_bf16_bits(w)
    hi_f = (hi.astype(np.uint32) << 16).view(np.float32)
    lo = to_bf16_bits((w - hi_f).astype(np.float32))
    return hi, lo
