"""keras.backend subset (numpy float64, eager)."""
import numpy as np

from .engine import KTensor, get_uid  # noqa: F401

_FORMAT = ['channels_last']


def _v(x):
    return x.value if isinstance(x, KTensor) else np.asarray(x, dtype=np.float64)


def _t(v):
    return KTensor(v)


def image_data_format():
    return _FORMAT[0]


def set_image_data_format(f):
    _FORMAT[0] = f


def epsilon():
    return 1e-7


def floatx():
    return 'float32'


def int_shape(x):
    if isinstance(x, KTensor):
        return x.shape
    return tuple(np.shape(x))


def ndim(x):
    return len(int_shape(x))


def shape(x):
    return int_shape(x)


def expand_dims(x, axis=-1):
    return _t(np.expand_dims(_v(x), axis))


def squeeze(x, axis):
    return _t(np.squeeze(_v(x), axis=axis))


def tile(x, n):
    return _t(np.tile(_v(x), n))


def reshape(x, shape):
    v = _v(x)
    shape = tuple(-1 if s is None else s for s in shape)
    return _t(v.reshape(shape))


def sum(x, axis=None, keepdims=False):
    if isinstance(axis, list):
        axis = tuple(axis)
    return _t(np.sum(_v(x), axis=axis, keepdims=keepdims))


def max(x, axis=None, keepdims=False):
    if isinstance(axis, list):
        axis = tuple(axis)
    return _t(np.max(_v(x), axis=axis, keepdims=keepdims))


def min(x, axis=None, keepdims=False):
    if isinstance(axis, list):
        axis = tuple(axis)
    return _t(np.min(_v(x), axis=axis, keepdims=keepdims))


def mean(x, axis=None, keepdims=False):
    if isinstance(axis, list):
        axis = tuple(axis)
    return _t(np.mean(_v(x), axis=axis, keepdims=keepdims))


def exp(x):
    return _t(np.exp(_v(x)))


def log(x):
    return _t(np.log(_v(x)))


def sqrt(x):
    return _t(np.sqrt(_v(x)))


def square(x):
    return _t(np.square(_v(x)))


def abs(x):
    return _t(np.abs(_v(x)))


def clip(x, lo, hi):
    return _t(np.clip(_v(x), lo, hi))


def cast(x, dtype):
    return _t(_v(x).astype(np.float64))


def greater_equal(x, y):
    return _t((_v(x) >= _v(y)).astype(np.float64))


def greater(x, y):
    return _t((_v(x) > _v(y)).astype(np.float64))


def stop_gradient(x):
    return x


def concatenate(xs, axis=-1):
    return _t(np.concatenate([_v(x) for x in xs], axis=axis))


def permute_dimensions(x, pattern):
    return _t(np.transpose(_v(x), pattern))


def softmax(x, axis=-1):
    v = _v(x)
    e = np.exp(v - v.max(axis=axis, keepdims=True))
    return _t(e / e.sum(axis=axis, keepdims=True))


def relu(x, alpha=0.0, max_value=None):
    v = _v(x)
    return _t(np.where(v > 0, v, alpha * v))


def sigmoid(x):
    return _t(1.0 / (1.0 + np.exp(-_v(x))))


def constant(value, dtype=None, shape=None, name=None):
    return _t(np.asarray(value, dtype=np.float64))


def variable(value, dtype=None, name=None):
    return _t(np.asarray(value, dtype=np.float64))


def clear_session():
    pass
