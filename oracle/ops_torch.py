"""torch-CPU (oneDNN) implementation of the same op set as ops_np.py.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Two jobs: (1) an independent cross-check of the numpy restatement, (2) the timed
"CPU port" baseline (`bench.py` cpu_baseline / --impl reference) -- the closest
runnable stand-in for the reference's Keras/TF CPU forward, which cannot run here
(no tensorflow / keras in the image).  Tensors are NHWC-shaped torch tensors;
convolutions see them as channels_last NCHW views (no copies).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops_np

EPS_BN = ops_np.EPS_BN
K_EPSILON = ops_np.K_EPSILON
DTYPE = torch.float32


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1)


def _pads(h, w, kh, kw, sh, sw, padding):
    ho, wo, pt, pb, pl, pr = ops_np._out_and_pad(h, w, kh, kw, sh, sw, padding)
    return pt, pb, pl, pr


def conv2d(x, w, strides=(1, 1), padding='same'):
    kh, kw, cin, cout = w.shape
    pt, pb, pl, pr = _pads(x.shape[1], x.shape[2], kh, kw, strides[0], strides[1], padding)
    xi = _nchw(x)
    if pt or pb or pl or pr:
        xi = F.pad(xi, (pl, pr, pt, pb))
    wt = w.permute(3, 2, 0, 1).contiguous(memory_format=torch.channels_last) if kh * kw > 1 \
        else w.permute(3, 2, 0, 1)
    return _nhwc(F.conv2d(xi, wt, stride=strides))


def depthwise_conv2d(x, w, strides=(1, 1), padding='same'):
    kh, kw, c, _ = w.shape
    pt, pb, pl, pr = _pads(x.shape[1], x.shape[2], kh, kw, strides[0], strides[1], padding)
    xi = _nchw(x)
    if pt or pb or pl or pr:
        xi = F.pad(xi, (pl, pr, pt, pb))
    wt = w.permute(2, 3, 0, 1)          # (C,1,kh,kw)
    return _nhwc(F.conv2d(xi, wt, stride=strides, groups=c))


def separable_conv2d(x, dw, pw, strides=(1, 1), padding='same'):
    return conv2d(depthwise_conv2d(x, dw, strides, padding), pw, (1, 1), 'valid')


def batchnorm(x, gamma, beta, mean, var):
    y = (x - mean) / torch.sqrt(var + EPS_BN)
    if gamma is not None:
        y = y * gamma
    return y + beta


def relu(x):
    return torch.relu(x)


def sigmoid(x):
    return torch.sigmoid(x)


def softmax(x):
    return torch.softmax(x, dim=-1)


def maxpool2d(x, pool=(2, 2), strides=None, padding='valid'):
    if strides is None:
        strides = pool
    pt, pb, pl, pr = _pads(x.shape[1], x.shape[2], pool[0], pool[1], strides[0], strides[1], padding)
    xi = _nchw(x)
    if pt or pb or pl or pr:
        xi = F.pad(xi, (pl, pr, pt, pb), value=float('-inf'))
    return _nhwc(F.max_pool2d(xi, pool, strides))


def avgpool2d_2x2_s1_valid(x):
    return 0.25 * (x[:, :-1, :-1] + x[:, :-1, 1:] + x[:, 1:, :-1] + x[:, 1:, 1:])


def global_maxpool2d(x):
    return x.amax(dim=(1, 2))


def upsample2d(x, size=(2, 2)):
    return x.repeat_interleave(size[0], dim=1).repeat_interleave(size[1], dim=2)


def zeropad2d(x, pads):
    (pt, pb), (pl, pr) = pads
    return F.pad(x, (0, 0, pl, pr, pt, pb))


def concat(ts):
    return torch.cat(ts, dim=-1)


def channel_softmax_2d(x, alpha=1):
    if alpha != 1:
        x = alpha * x
    e = torch.exp(x - x.amax(dim=(-3, -2), keepdim=True))
    s = torch.clamp(e.sum(dim=(-3, -2), keepdim=True), min=K_EPSILON)
    return e / s


def channel_softmax_1d(x):
    e = torch.exp(x - x.amax(dim=1, keepdim=True))
    return e / e.sum(dim=1, keepdim=True)


def lin_interpolation_2d(p, axis):
    grid = torch.from_numpy(ops_np.linspace_2d(p.shape[1], p.shape[2], dim=axis)).to(p.dtype)
    return torch.einsum('nrqc,rq->nc', p, grid)[..., None]


def softargmax2d(p):
    return torch.cat([lin_interpolation_2d(p, 0), lin_interpolation_2d(p, 1)], dim=-1)


def lin_interpolation_1d(p):
    d = p.shape[1]
    start = 1 / (2 * d)
    lin = torch.from_numpy(np.linspace(start, 1 - start, num=d).astype(np.float32)).to(p.dtype)
    return torch.einsum('ndc,d->nc', p, lin)[..., None]


def keypoint_confidence(p):
    return global_maxpool2d(4 * avgpool2d_2x2_s1_valid(p))[..., None]


def max_min_pooling(x, pool=(2, 2), padding='same'):
    return maxpool2d(x, pool, None, padding) - maxpool2d(-x, pool, None, padding)


def global_max_min_pooling(x):
    return global_maxpool2d(x) - global_maxpool2d(-x)


def kronecker_prod(h, f):
    return torch.einsum('bthwj,bthwf->btjf', h, f)


def time_distributed(fn, x):
    b, t = x.shape[:2]
    y = fn(x.reshape((b * t,) + tuple(x.shape[2:])))
    return y.reshape((b, t) + tuple(y.shape[1:]))


def from_numpy(a):
    if isinstance(a, torch.Tensor):
        return a.to(DTYPE)
    return torch.from_numpy(np.ascontiguousarray(a)).to(DTYPE)


def to_numpy(a):
    return a.detach().cpu().numpy()


def mean(x, axes):
    return x.mean(dim=axes)


def amax(x, axes):
    return x.amax(dim=axes)


def asum(x, axes):
    return x.sum(dim=axes)
