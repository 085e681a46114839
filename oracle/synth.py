"""Seeded synthetic weights / frames for the oracle.
TEST INFRASTRUCTURE (see oracle/__init__.py).

The recipe is restated independently in deephar_b200/weights.py; tests assert that
both produce identical arrays for identical (seed, name, shape, calibration).

Calibration: purely random BatchNorm statistics let the activations of the 8
stacked blocks grow ~13x per block (measured), which saturates every soft-argmax
into a one-hot -- a degenerate test.  `Calibrator` therefore runs the oracle once
and records, per BatchNorm layer, the scalar mean / variance of its input (so the
synthetic moving statistics normalise like trained ones do), and per heat-map head
conv a scalar gain that gives the logits a standard deviation of `HEAD_STD`.  The
resulting small table (two scalars per layer) is committed under
deephar_b200/synth_calib/ and consumed by both generators.
"""
import zlib

import numpy as np

HEAD_STD = 3.0
HEAD_MARKERS = ('RegMap', '_heatmaps_conv1', '_depthmaps_conv1', '_pred_conv2h')


def is_head_kernel(name):
    return name.endswith('/kernel') and any(m in name for m in HEAD_MARKERS)


def synth_weight(seed, name, shape, calib=None):
    calib = calib or {}
    rng = np.random.default_rng([int(seed), zlib.crc32(name.encode('utf-8'))])
    layer, leaf = name.rsplit('/', 1)
    if leaf == 'kernel':                       # Conv2D (kh,kw,Cin,Cout): N(0, 2/fan_in)
        fan_in = shape[0] * shape[1] * shape[2]
        a = rng.standard_normal(shape) * np.sqrt(2.0 / fan_in) * calib.get(name, 1.0)
    elif leaf == 'pointwise_kernel':           # (1,1,Cin,Cout): N(0, 2/Cin)
        a = rng.standard_normal(shape) * np.sqrt(2.0 / shape[2])
    elif leaf == 'depthwise_kernel':           # (kh,kw,Cin,1): N(0, 1/(kh*kw))
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


class SyntheticTable(object):
    """dict-like: generates each weight on first lookup (needs the shape)."""

    def __init__(self, seed=1234, calib=None):
        self.seed = seed
        self.calib = dict(calib or {})
        self.cache = {}

    def lookup(self, name, shape):
        if name not in self.cache:
            self.cache[name] = synth_weight(self.seed, name, tuple(shape), self.calib)
        return self.cache[name]


class Calibrator(SyntheticTable):
    """SyntheticTable that fills `calib` while the oracle runs (see module doc)."""

    def observe_bn(self, layer, x):
        x = np.asarray(x, dtype=np.float64)
        red = tuple(range(x.ndim - 1))
        m = float(x.mean(axis=red).mean())
        v = float(x.var(axis=red).mean())
        self.calib[layer] = (round(m, 6), round(max(v, 1e-6), 6))

    def observe_head(self, name, y_unit_gain):
        s = float(np.asarray(y_unit_gain, dtype=np.float64).std())
        g = round(HEAD_STD / max(s, 1e-12), 6)
        self.calib[name] = g
        self.cache[name] = (self.cache[name] * np.float32(1.0)).astype(np.float32)
        # regenerate with the recorded (rounded) gain so that it matches synth_weight
        self.cache[name] = synth_weight(self.seed, name, self.cache[name].shape, self.calib)
        return g


def synth_frames(n, h=256, w=256, seed=0):
    """uniform [-1,1] fp32 NHWC frames (range of utils/transform.py:212-231)."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, (n, h, w, 3)).astype(np.float32)
