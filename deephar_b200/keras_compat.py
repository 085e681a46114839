"""Keras-signature front end: write (or keep) model code in the `keras.layers` functional style the reference's
builders use -- `Conv2D(filters, size, use_bias=False, name=...)(x)`, `BatchNormalization(name=...)(x)`,
`add([a, b])`, `Model(inputs=inp, outputs=outs)` -- and get a compiled B200 model instead of a TensorFlow graph.

    from deephar_b200.keras_compat import Input, Conv2D, BatchNormalization, Activation, add, Model
    inp = Input(shape=(256, 256, 3))
    x = Conv2D(64, (7, 7), strides=(2, 2), padding='same', use_bias=False, name='conv1')(inp)
    x = Activation('relu')(BatchNormalization(name='bn1')(x))
    model = Model(inputs=inp, outputs=[x])          # -> deephar_b200.model.Model: predict / load_weights / ...

Every layer object records a node of the symbolic graph (graph.py) when it is called; nothing is computed here.
Layer classes, argument names and DEFAULTS are those of Keras 2.1.4 (the version the reference pins: Conv2D padding
'valid', use_bias True, BatchNormalization scale/center True, MaxPooling2D strides = pool_size ...), so code written for
the reference means the same thing.  What the forward path of the reference models never uses is rejected at build
time, not silently approximated: biases (all reference convs are `use_bias=False`, deephar/layers.py:69,78),
activations fused into a conv via `activation=`, dilation, depth_multiplier != 1, `data_format='channels_first'`.
The reference's own helper layer on top of Keras (deephar/layers.py, deephar/activations.py) is mirrored function-style
in deephar_b200.layers and re-exported here (`channel_softmax_2d`, `softargmax2d`, `kronecker_prod`, ...).
`TimeDistributed(layer, name=...)` is the identity on the folded B*T frame axis and passes its name to the wrapped
layer, which is how Keras names the weights of a wrapped layer in a checkpoint.
"""
from . import layers as L
from .graph import Graph
from .layers import (act_channel_softmax, channel_slice, channel_softmax_2d, depth_expectation,  # noqa: F401
                     global_max_min_pooling, keypoint_confidence, kronecker_prod, max_min_pooling, softargmax2d)
from .model import Model as _Model


def Input(shape=None, batch_shape=None, name=None, graph=None, frames_per_clip=None):
    """keras.layers.Input.  shape (H, W, C) -> per-frame tensor; shape (T, H, W, C) -> a clip input whose frames are
    folded into the batch axis (what TimeDistributed does in the reference, deephar/models/spnet.py:283-296).  A new
    symbolic graph starts at every Input unless `graph` is given."""
    if shape is None and batch_shape is not None:
        shape = tuple(batch_shape)[1:]
    if shape is None:
        raise ValueError('Input: shape is required')
    shape = tuple(int(s) for s in shape)
    g = graph if graph is not None else Graph(name or 'model')
    if len(shape) == 4:
        g.frames_per_clip = int(shape[0])
        shape = shape[1:]
    elif frames_per_clip:
        g.frames_per_clip = int(frames_per_clip)
    if len(shape) != 3:
        raise ValueError('Input: expected (H, W, C) or (T, H, W, C), got %r' % (shape,))
    return g.input(shape)


class Layer(object):
    """A Keras layer object: constructed with its hyper-parameters, applied by calling it on a tensor."""

    def __init__(self, name=None, **kwargs):
        unknown = set(kwargs) - {'trainable', 'input_shape', 'dtype'}
        if unknown:
            raise TypeError('%s: unsupported argument(s) %s' % (type(self).__name__, sorted(unknown)))
        self.name = name
        self._built = False

    def __call__(self, x):
        if self._built and self._has_weights:
            raise NotImplementedError('%s %r: calling a weighted layer twice (weight sharing) is not used by the '
                                      'reference models' % (type(self).__name__, self.name))
        self._built = True
        return self.call(x)

    _has_weights = False


def _no(cond, what):
    if cond:
        raise NotImplementedError(what + ' is not on the forward path of the reference models')


class Conv2D(Layer):
    _has_weights = True

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', data_format=None, dilation_rate=(1, 1),
                 activation=None, use_bias=True, name=None, **kwargs):
        Layer.__init__(self, name, **{k: v for k, v in kwargs.items() if not k.endswith(('_initializer', '_regularizer',
                                                                                           '_constraint'))})
        _no(use_bias, 'Conv2D(use_bias=True)')
        _no(activation not in (None, 'linear'), 'Conv2D(activation=...)')
        _no(L._pair(dilation_rate) != (1, 1), 'dilated convolution')
        _no(data_format not in (None, 'channels_last'), 'channels_first')
        self.filters, self.kernel_size, self.strides, self.padding = int(filters), kernel_size, strides, padding

    def call(self, x):
        return L.conv2d(x, self.filters, self.kernel_size, self.strides, self.padding, name=self.name)


class SeparableConv2D(Layer):
    _has_weights = True

    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', data_format=None, depth_multiplier=1,
                 activation=None, use_bias=True, name=None, **kwargs):
        Layer.__init__(self, name, **{k: v for k, v in kwargs.items() if not k.endswith(('_initializer', '_regularizer',
                                                                                           '_constraint'))})
        _no(use_bias, 'SeparableConv2D(use_bias=True)')
        _no(activation not in (None, 'linear'), 'SeparableConv2D(activation=...)')
        _no(depth_multiplier != 1, 'depth_multiplier != 1')
        _no(data_format not in (None, 'channels_last'), 'channels_first')
        self.filters, self.kernel_size, self.strides, self.padding = int(filters), kernel_size, strides, padding

    def call(self, x):
        return L.sepconv2d(x, self.filters, self.kernel_size, self.strides, self.padding, name=self.name)


class BatchNormalization(Layer):
    _has_weights = True

    def __init__(self, axis=-1, momentum=0.99, epsilon=1e-3, center=True, scale=True, name=None, **kwargs):
        Layer.__init__(self, name, **{k: v for k, v in kwargs.items() if not k.endswith(('_initializer', '_regularizer',
                                                                                           '_constraint'))})
        _no(axis not in (-1, 3), 'BatchNormalization over an axis other than channels')
        _no(not center, 'BatchNormalization(center=False)')
        if abs(float(epsilon) - 1e-3) > 1e-12:
            raise NotImplementedError('BatchNormalization: epsilon is fixed at the Keras default 1e-3 (got %r)' % epsilon)
        self.scale = bool(scale)

    def call(self, x):
        return L.BatchNormalization(x, scale=self.scale, name=self.name)


class Activation(Layer):
    def __init__(self, activation, name=None, **kwargs):
        Layer.__init__(self, name, **kwargs)
        if activation not in ('relu', 'sigmoid', 'softmax'):
            raise NotImplementedError('Activation(%r): the reference models use relu / sigmoid / softmax and the '
                                      'helpers channel_softmax_2d(...)' % (activation,))
        self.activation = activation

    def call(self, x):
        if self.activation == 'softmax' and x.kind != 'frame':
            return L.softmax_lastaxis(x, name=self.name)
        return L.Activation(x, self.activation, name=self.name)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding='valid', data_format=None, name=None, **kwargs):
        Layer.__init__(self, name, **kwargs)
        self.pool_size, self.strides, self.padding = pool_size, strides, padding

    def call(self, x):
        return L.MaxPooling2D(x, self.pool_size, self.strides, self.padding, name=self.name)


class UpSampling2D(Layer):
    def __init__(self, size=(2, 2), data_format=None, name=None, **kwargs):
        Layer.__init__(self, name, **kwargs)
        self.size = size

    def call(self, x):
        return L.UpSampling2D(x, self.size, name=self.name)


class ZeroPadding2D(Layer):
    def __init__(self, padding=(1, 1), data_format=None, name=None, **kwargs):
        Layer.__init__(self, name, **kwargs)
        if isinstance(padding, int):
            padding = ((padding, padding), (padding, padding))
        elif isinstance(padding[0], int):
            padding = ((padding[0], padding[0]), (padding[1], padding[1]))
        self.padding = tuple(tuple(p) for p in padding)

    def call(self, x):
        return L.ZeroPadding2D(x, self.padding)


class TimeDistributed(Layer):
    """keras.layers.TimeDistributed: frames already sit on the batch axis, so the wrapper only hands its name down."""

    def __init__(self, layer, name=None, **kwargs):
        Layer.__init__(self, name, **kwargs)
        self.layer = layer
        if name is not None:
            layer.name = name

    def call(self, x):
        return self.layer(x)


class _Merge(Layer):
    fn = None

    def call(self, ts):
        return type(self).fn(list(ts), name=self.name)


class Add(_Merge):
    fn = staticmethod(L.add)


class Concatenate(_Merge):
    fn = staticmethod(L.concatenate)

    def __init__(self, axis=-1, name=None, **kwargs):
        _Merge.__init__(self, name, **kwargs)
        _no(axis != -1, 'concatenation over an axis other than channels')


class Multiply(_Merge):
    fn = staticmethod(L.multiply)


def add(inputs, name=None):
    return Add(name=name)(inputs)


def concatenate(inputs, axis=-1, name=None):
    return Concatenate(axis=axis, name=name)(inputs)


def multiply(inputs, name=None):
    return Multiply(name=name)(inputs)


def Model(inputs=None, outputs=None, name=None):
    """keras.models.Model(inputs, outputs): compile the recorded graph for B200.  Layers that do not reach an output
    are dropped, as Keras drops them (their weights become optional when loading a checkpoint)."""
    inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
    outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
    if len(inputs) != 1:
        raise NotImplementedError('Model: exactly one Input (the frame / clip tensor) is supported')
    g = inputs[0].g
    for t in outputs:
        if t.g is not g:
            raise ValueError('Model: an output does not descend from the given Input')
    if g.inputs != inputs:
        raise ValueError('Model: `inputs` must be the Input the graph was started from')
    g.outputs = outputs
    if name:
        g.name = name
    return _Model(g, name=name)
