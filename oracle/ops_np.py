"""numpy restatement of the Keras 2.1.4 / TF 1.6 (channels_last) ops used by the
deephar forward path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

All tensors are NHWC.  The dtype of the input is kept (the oracle runs fp64).
Semantics follow SURVEY.md Appendix A; each function names the reference call
sites it stands in for.
"""
import math

import numpy as np

EPS_BN = 1e-3      # keras BatchNormalization default epsilon
K_EPSILON = 1e-7   # keras.backend.epsilon()


# ----------------------------------------------------------------------------
# padding helpers (TF "SAME": extra pad goes to bottom/right)
# ----------------------------------------------------------------------------
def same_pad(in_size, k, s):
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def _out_and_pad(h, w, kh, kw, sh, sw, padding):
    if padding == 'same':
        ho, pt, pb = same_pad(h, kh, sh)
        wo, pl, pr = same_pad(w, kw, sw)
    elif padding == 'valid':
        ho, wo = (h - kh) // sh + 1, (w - kw) // sw + 1
        pt = pb = pl = pr = 0
    else:
        raise ValueError(padding)
    return ho, wo, pt, pb, pl, pr


def _pad(x, pt, pb, pl, pr, value=0.0):
    if pt == pb == pl == pr == 0:
        return x
    return np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)), mode='constant',
                  constant_values=value)


# ----------------------------------------------------------------------------
# convolutions (layers.py:66-80 conv2d / sepconv2d; use_bias=False everywhere)
# ----------------------------------------------------------------------------
def conv2d(x, w, strides=(1, 1), padding='same'):
    """keras Conv2D == tf.nn.conv2d: cross-correlation, kernel (kh,kw,Cin,Cout)."""
    n, h, wd, cin = x.shape
    kh, kw, cin2, cout = w.shape
    assert cin == cin2, (x.shape, w.shape)
    sh, sw = strides
    ho, wo, pt, pb, pl, pr = _out_and_pad(h, wd, kh, kw, sh, sw, padding)
    xp = _pad(x, pt, pb, pl, pr)
    out = np.zeros((n * ho * wo, cout), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            xs = xp[:, i:i + (ho - 1) * sh + 1:sh, j:j + (wo - 1) * sw + 1:sw, :]
            out += xs.reshape(-1, cin) @ w[i, j].astype(x.dtype)
    return out.reshape(n, ho, wo, cout)


def depthwise_conv2d(x, w, strides=(1, 1), padding='same'):
    """tf.nn.depthwise_conv2d with kernel (kh,kw,Cin,1)."""
    n, h, wd, c = x.shape
    kh, kw, c2, mult = w.shape
    assert c == c2 and mult == 1
    sh, sw = strides
    ho, wo, pt, pb, pl, pr = _out_and_pad(h, wd, kh, kw, sh, sw, padding)
    xp = _pad(x, pt, pb, pl, pr)
    out = np.zeros((n, ho, wo, c), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            xs = xp[:, i:i + (ho - 1) * sh + 1:sh, j:j + (wo - 1) * sw + 1:sw, :]
            out += xs * w[i, j, :, 0].astype(x.dtype)
    return out


def separable_conv2d(x, dw, pw, strides=(1, 1), padding='same'):
    """keras SeparableConv2D == tf.nn.separable_conv2d: depthwise (stride here)
    then 1x1 pointwise, nothing in between."""
    y = depthwise_conv2d(x, dw, strides, padding)
    return conv2d(y, pw, (1, 1), 'valid')


def batchnorm(x, gamma, beta, mean, var):
    """keras BatchNormalization(axis=-1) at inference; gamma may be None (scale=False)."""
    y = (x - mean) / np.sqrt(var + EPS_BN)
    if gamma is not None:
        y = y * gamma
    return y + beta


def relu(x):
    return np.maximum(x, 0)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def softmax(x):
    """keras Activation('softmax'): last axis."""
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return e / e.sum(axis=-1, keepdims=True)


# ----------------------------------------------------------------------------
# pooling / resampling
# ----------------------------------------------------------------------------
def maxpool2d(x, pool=(2, 2), strides=None, padding='valid'):
    """keras MaxPooling2D: strides default to pool; 'same' pads with -inf."""
    if strides is None:
        strides = pool
    n, h, wd, c = x.shape
    kh, kw = pool
    sh, sw = strides
    ho, wo, pt, pb, pl, pr = _out_and_pad(h, wd, kh, kw, sh, sw, padding)
    xp = _pad(x, pt, pb, pl, pr, value=-np.inf)
    out = np.full((n, ho, wo, c), -np.inf, dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            xs = xp[:, i:i + (ho - 1) * sh + 1:sh, j:j + (wo - 1) * sw + 1:sw, :]
            out = np.maximum(out, xs)
    return out


def avgpool2d_2x2_s1_valid(x):
    """keras AveragePooling2D((2,2), strides=(1,1)) -- default padding 'valid'."""
    return 0.25 * (x[:, :-1, :-1] + x[:, :-1, 1:] + x[:, 1:, :-1] + x[:, 1:, 1:])


def global_maxpool2d(x):
    return x.max(axis=(1, 2))


def upsample2d(x, size=(2, 2)):
    """keras UpSampling2D: nearest-neighbour repeat."""
    return np.repeat(np.repeat(x, size[0], axis=1), size[1], axis=2)


def zeropad2d(x, pads):
    (pt, pb), (pl, pr) = pads
    return _pad(x, pt, pb, pl, pr)


def concat(ts):
    return np.concatenate(ts, axis=-1)


# ----------------------------------------------------------------------------
# deephar/activations.py
# ----------------------------------------------------------------------------
def channel_softmax_2d(x, alpha=1):
    """activations.py:3-16 -- softmax over the two spatial axes (-3,-2)."""
    assert x.ndim in (4, 5)
    if alpha != 1:
        x = alpha * x
    e = np.exp(x - x.max(axis=(-3, -2), keepdims=True))
    s = np.clip(e.sum(axis=(-3, -2), keepdims=True), K_EPSILON, None)
    return e / s


def channel_softmax_1d(x):
    """activations.py:18-30 -- softmax over axis 1 of a (N, D, C) tensor."""
    assert x.ndim == 3
    e = np.exp(x - x.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


# ----------------------------------------------------------------------------
# deephar/utils/math.py:6-19 and the fixed-weight "soft-argmax as convolution"
# layers (layers.py:122-200)
# ----------------------------------------------------------------------------
def linspace_2d(nb_rows, nb_cols, dim=0):
    """utils/math.py:6-19.  dim=0: every row is linspace(0,1,nb_cols) (x grid);
    dim=1: every column is linspace(0,1,nb_rows) (y grid).  float32 in the
    reference (the grid becomes float32 conv weights)."""
    if dim == 1:
        col = np.linspace(0.0, 1.0, num=nb_rows).astype(np.float32)
        return np.repeat(col[:, None], nb_cols, axis=1)
    row = np.linspace(0.0, 1.0, num=nb_cols).astype(np.float32)
    return np.repeat(row[None, :], nb_rows, axis=0)


def lin_interpolation_2d(p, axis):
    """layers.py:160-200: SeparableConv2D(C,(H,W),valid) with depthwise = grid and
    pointwise = I, i.e. out[n,c] = sum_{r,q} p[n,r,q,c] * grid[r,q].  -> (N,C,1).
    The vmin/vmax arguments of the reference are ignored there too (App. C.1)."""
    n, h, w, c = p.shape
    grid = linspace_2d(h, w, dim=axis).astype(p.dtype)
    return np.einsum('nrqc,rq->nc', p, grid)[..., None]


def softargmax2d(p):
    """layers.py:122-129: concat(E[x], E[y]) -> (N,C,2)."""
    return np.concatenate([lin_interpolation_2d(p, 0), lin_interpolation_2d(p, 1)], axis=-1)


def lin_interpolation_1d(p):
    """layers.py:132-157: Conv1D(C, D, valid) with w[:, i, i] = linspace(1/2D, 1-1/2D, D).
    p: (N, D, C) -> (N, C, 1)."""
    n, d, c = p.shape
    start = 1 / (2 * d)
    lin = np.linspace(start, 1 - start, num=d).astype(np.float32).astype(p.dtype)
    return np.einsum('ndc,d->nc', p, lin)[..., None]


def keypoint_confidence(p):
    """layers.py:107-119: 4*avgpool2x2(s1,valid) then global max -> (N,C,1)."""
    return global_maxpool2d(4 * avgpool2d_2x2_s1_valid(p))[..., None]


def max_min_pooling(x, pool=(2, 2), padding='same'):
    """layers.py:411-425: maxpool(x) - maxpool(-x) (strides default to pool)."""
    return maxpool2d(x, pool, None, padding) - maxpool2d(-x, pool, None, padding)


def global_max_min_pooling(x):
    """layers.py:428-442."""
    return global_maxpool2d(x) - global_maxpool2d(-x)


def kronecker_prod(h, f):
    """layers.py:478-508 for clip tensors: h (B,T,H,W,nj), f (B,T,H,W,F) ->
    (B,T,nj,F) = sum_{hw} h * f  (sum over axes (2,3))."""
    assert h.ndim == 5 and f.ndim == 5
    return np.einsum('bthwj,bthwf->btjf', h, f)


def time_distributed(fn, x):
    """keras TimeDistributed: fold (B,T,...) -> (B*T,...), apply, unfold."""
    b, t = x.shape[:2]
    y = fn(x.reshape((b * t,) + x.shape[2:]))
    return y.reshape((b, t) + y.shape[1:])


def from_numpy(a, dtype=np.float64):
    return np.asarray(a, dtype=dtype)


def to_numpy(a):
    return np.asarray(a)


def mean(x, axes):
    return x.mean(axis=axes)


def amax(x, axes):
    return x.max(axis=axes)


def asum(x, axes):
    return x.sum(axis=axes)
