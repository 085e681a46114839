"""keras.layers subset, eager float64.  Convolutions run on torch-CPU float64 (independent of oracle/ops_np.py);
TF 'SAME' padding: total = max((ceil(in / s) - 1) * s + k - in, 0), before = total // 2 (extra pixel after)."""
import numpy as np
import torch
import torch.nn.functional as F

from .engine import Input, InputLayer, KTensor, Layer, Model, get_uid  # noqa: F401


def _pair(v, n=2):
    if isinstance(v, (list, tuple)):
        return tuple(int(a) for a in v)
    return (int(v),) * n


def _same_pads(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _pad_hw(x, k, s, padding, value=0.0):
    """x: (N,H,W,C) numpy -> padded copy."""
    if padding == 'valid':
        return x
    assert padding == 'same', padding
    (pt, pb), (pl, pr) = _same_pads(x.shape[1], k[0], s[0]), _same_pads(x.shape[2], k[1], s[1])
    return np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)), mode='constant', constant_values=value)


def _conv2d(x, w, strides, padding, groups=1):
    """x (N,H,W,C), w (kh,kw,Cin/groups,Cout) -> (N,Ho,Wo,Cout)."""
    xp = _pad_hw(x, w.shape[:2], strides, padding)
    xt = torch.from_numpy(np.ascontiguousarray(xp.transpose(0, 3, 1, 2)))
    wt = torch.from_numpy(np.ascontiguousarray(w.transpose(3, 2, 0, 1)))
    y = F.conv2d(xt, wt, stride=strides, groups=groups)
    return y.numpy().transpose(0, 2, 3, 1)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', use_bias=True, activation=None,
                 data_format=None, **kw):
        super(Conv2D, self).__init__(**kw)
        self.filters, self.kernel_size, self.strides = int(filters), _pair(kernel_size), _pair(strides)
        self.padding, self.use_bias = padding, use_bias
        assert activation is None

    def build(self, s):
        self.kernel = self.add_weight('kernel', self.kernel_size + (s[-1], self.filters))
        self.bias = self.add_weight('bias', (self.filters,)) if self.use_bias else None

    def call(self, x):
        y = _conv2d(x, self.kernel['value'], self.strides, self.padding)
        if self.bias is not None:
            y = y + self.bias['value']
        return y


class SeparableConv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', use_bias=True, depth_multiplier=1, **kw):
        super(SeparableConv2D, self).__init__(**kw)
        assert depth_multiplier == 1
        self.filters, self.kernel_size, self.strides = int(filters), _pair(kernel_size), _pair(strides)
        self.padding, self.use_bias = padding, use_bias

    def build(self, s):
        self.dw = self.add_weight('depthwise_kernel', self.kernel_size + (s[-1], 1))
        self.pw = self.add_weight('pointwise_kernel', (1, 1, s[-1], self.filters))
        self.bias = self.add_weight('bias', (self.filters,)) if self.use_bias else None

    def call(self, x):
        c = x.shape[-1]
        wd = self.dw['value'].transpose(0, 1, 3, 2)                # (kh,kw,1,C): one filter per group
        y = _conv2d(x, wd, self.strides, self.padding, groups=c)
        y = _conv2d(y, self.pw['value'], (1, 1), 'valid')
        if self.bias is not None:
            y = y + self.bias['value']
        return y


class Conv1D(Layer):
    def __init__(self, filters, kernel_size, strides=1, padding='valid', use_bias=True, **kw):
        super(Conv1D, self).__init__(**kw)
        self.filters, self.k, self.s, self.padding, self.use_bias = int(filters), int(kernel_size), int(strides), padding, use_bias

    def build(self, s):
        self.kernel = self.add_weight('kernel', (self.k, s[-1], self.filters))
        self.bias = self.add_weight('bias', (self.filters,)) if self.use_bias else None

    def call(self, x):
        y = _conv2d(x[:, :, None, :], self.kernel['value'][:, None, :, :], (self.s, 1), self.padding)[:, :, 0, :]
        if self.bias is not None:
            y = y + self.bias['value']
        return y


class Dense(Layer):
    def __init__(self, units, use_bias=True, activation=None, **kw):
        super(Dense, self).__init__(**kw)
        self.units, self.use_bias = int(units), use_bias
        assert activation is None

    def build(self, s):
        self.kernel = self.add_weight('kernel', (s[-1], self.units))
        self.bias = self.add_weight('bias', (self.units,)) if self.use_bias else None

    def call(self, x):
        y = x @ self.kernel['value']
        if self.bias is not None:
            y = y + self.bias['value']
        return y


class BatchNormalization(Layer):
    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, **kw):
        super(BatchNormalization, self).__init__(**kw)
        assert axis == -1
        self.eps, self.center, self.scale = epsilon, center, scale

    def build(self, s):
        c = s[-1]
        self.gamma = self.add_weight('gamma', (c,), init=1.0) if self.scale else None
        self.beta = self.add_weight('beta', (c,)) if self.center else None
        self.mean = self.add_weight('moving_mean', (c,), trainable=False)
        self.var = self.add_weight('moving_variance', (c,), trainable=False, init=1.0)

    def call(self, x):
        y = (x - self.mean['value']) / np.sqrt(self.var['value'] + self.eps)
        if self.gamma is not None:
            y = y * self.gamma['value']
        if self.beta is not None:
            y = y + self.beta['value']
        return y


def _act(name, x):
    if name == 'relu':
        return np.maximum(x, 0.0)
    if name == 'sigmoid':
        return 1.0 / (1.0 + np.exp(-x))
    if name == 'tanh':
        return np.tanh(x)
    if name == 'softmax':
        e = np.exp(x - x.max(axis=-1, keepdims=True))
        return e / e.sum(axis=-1, keepdims=True)
    if name in (None, 'linear'):
        return x
    raise NotImplementedError(name)


class Activation(Layer):
    def __init__(self, activation, **kw):
        super(Activation, self).__init__(**kw)
        self.activation = activation

    def call(self, x):
        if callable(self.activation):
            return self.activation(KTensor(x)).value
        return _act(self.activation, x)


class LeakyReLU(Layer):
    def __init__(self, alpha=0.3, **kw):
        super(LeakyReLU, self).__init__(**kw)
        self.alpha = alpha

    def call(self, x):
        return np.where(x > 0, x, self.alpha * x)


class Dropout(Layer):
    def __init__(self, rate, **kw):
        super(Dropout, self).__init__(**kw)

    def call(self, x):
        return x


class Flatten(Layer):
    def call(self, x):
        return x.reshape(x.shape[0], -1)


class Reshape(Layer):
    def __init__(self, target_shape, **kw):
        super(Reshape, self).__init__(**kw)
        self.target_shape = tuple(target_shape)

    def call(self, x):
        return x.reshape((x.shape[0],) + self.target_shape)


class Lambda(Layer):
    def __init__(self, function, output_shape=None, mask=None, arguments=None, **kw):
        super(Lambda, self).__init__(**kw)
        self.function, self.arguments = function, arguments or {}

    def call(self, x):
        arg = [KTensor(v) for v in x] if isinstance(x, (list, tuple)) else KTensor(x)
        out = self.function(arg, **self.arguments)
        if isinstance(out, (list, tuple)):
            return [o.value for o in out]
        return out.value


class _Pool2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', **kw):
        super(_Pool2D, self).__init__(**kw)
        self.pool_size = _pair(pool_size)
        self.strides = _pair(strides) if strides is not None else self.pool_size
        self.padding = padding

    def _windows(self, xp):
        kh, kw = self.pool_size
        sh, sw = self.strides
        ho = (xp.shape[1] - kh) // sh + 1
        wo = (xp.shape[2] - kw) // sw + 1
        return [xp[:, a:a + (ho - 1) * sh + 1:sh, b:b + (wo - 1) * sw + 1:sw, :] for a in range(kh) for b in range(kw)]


class MaxPooling2D(_Pool2D):
    def call(self, x):
        xp = _pad_hw(x, self.pool_size, self.strides, self.padding, value=-np.inf)
        return np.max(np.stack(self._windows(xp), 0), 0)


class AveragePooling2D(_Pool2D):
    def call(self, x):
        xp = _pad_hw(x, self.pool_size, self.strides, self.padding, value=0.0)
        s = np.sum(np.stack(self._windows(xp), 0), 0)
        ones = _pad_hw(np.ones_like(x), self.pool_size, self.strides, self.padding, value=0.0)
        n = np.sum(np.stack(self._windows(ones), 0), 0)          # TF: padded cells are not counted
        return s / n


class GlobalMaxPooling2D(Layer):
    def call(self, x):
        return x.max(axis=(1, 2))


class GlobalAveragePooling2D(Layer):
    def call(self, x):
        return x.mean(axis=(1, 2))


class GlobalMaxPooling1D(Layer):
    def call(self, x):
        return x.max(axis=1)


class GlobalAveragePooling1D(Layer):
    def call(self, x):
        return x.mean(axis=1)


class UpSampling2D(Layer):
    def __init__(self, size=(2, 2), **kw):
        super(UpSampling2D, self).__init__(**kw)
        self.size = _pair(size)

    def call(self, x):
        return np.repeat(np.repeat(x, self.size[0], axis=1), self.size[1], axis=2)


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), **kw):
        super(ZeroPadding2D, self).__init__(**kw)
        if isinstance(padding, int):
            padding = ((padding, padding), (padding, padding))
        elif isinstance(padding[0], int):
            padding = ((padding[0], padding[0]), (padding[1], padding[1]))
        self.padding = tuple(tuple(p) for p in padding)

    def call(self, x):
        return np.pad(x, ((0, 0), self.padding[0], self.padding[1], (0, 0)), mode='constant')


class TimeDistributed(Layer):
    """Applies `layer` to every temporal slice: (B,T,...) -> fold to (B*T,...) -> unfold."""

    def __init__(self, layer, **kw):
        super(TimeDistributed, self).__init__(**kw)
        self.layer = layer

    @property
    def weights(self):
        return self.layer.weights

    def get_weights(self):
        return self.layer.get_weights()

    def set_weights(self, v):
        self.layer.set_weights(v)

    def compute(self, vals, was_list):
        assert not was_list
        x = vals[0]
        b, t = x.shape[:2]
        y = self.layer.compute([x.reshape((b * t,) + x.shape[2:])], False)
        self.built = True
        return y.reshape((b, t) + y.shape[1:])


class _Merge(Layer):
    def __init__(self, **kw):
        super(_Merge, self).__init__(**kw)


class Add(_Merge):
    def call(self, xs):
        y = xs[0]
        for x in xs[1:]:
            y = y + x
        return y


class Multiply(_Merge):
    def call(self, xs):
        y = xs[0]
        for x in xs[1:]:
            y = y * x
        return y


class Average(_Merge):
    def call(self, xs):
        return Add.call(self, xs) / len(xs)


class Maximum(_Merge):
    def call(self, xs):
        y = xs[0]
        for x in xs[1:]:
            y = np.maximum(y, x)
        return y


class Concatenate(_Merge):
    def __init__(self, axis=-1, **kw):
        super(Concatenate, self).__init__(**kw)
        self.axis = axis

    def call(self, xs):
        return np.concatenate(xs, axis=self.axis)


def add(inputs, **kw):
    return Add(**kw)(inputs)


def multiply(inputs, **kw):
    return Multiply(**kw)(inputs)


def average(inputs, **kw):
    return Average(**kw)(inputs)


def maximum(inputs, **kw):
    return Maximum(**kw)(inputs)


def concatenate(inputs, axis=-1, **kw):
    return Concatenate(axis=axis, **kw)(inputs)


class _NotOnThePath(Layer):
    def __init__(self, *a, **k):
        raise NotImplementedError('keras_shim: %s is not used by the forward paths in scope' % self.__class__.__name__)


class Conv3D(_NotOnThePath): pass            # noqa: E701
class Conv2DTranspose(_NotOnThePath): pass   # noqa: E701
class LocallyConnected1D(_NotOnThePath): pass  # noqa: E701
class SimpleRNN(_NotOnThePath): pass         # noqa: E701
class LSTM(_NotOnThePath): pass              # noqa: E701
class MaxPooling3D(_NotOnThePath): pass      # noqa: E701
class GlobalMaxPooling3D(_NotOnThePath): pass  # noqa: E701
class UpSampling3D(_NotOnThePath): pass      # noqa: E701
