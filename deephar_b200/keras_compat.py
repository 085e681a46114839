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


# Keras numbers auto-named layers (conv2d_7, batch_normalization_12) with ONE counter per class for the whole
# session, across sub-models: every graph started here shares this table until clear_session().
_SESSION_COUNTERS = {}


def clear_session():
    """keras.backend.clear_session(): restart the auto-name counters (call between two independent models)."""
    _SESSION_COUNTERS.clear()


def Input(shape=None, batch_shape=None, name=None, graph=None, frames_per_clip=None):
    """keras.layers.Input.  shape (H, W, C) -> per-frame tensor; shape (T, H, W, C) -> a clip input whose frames are
    folded into the batch axis (what TimeDistributed does in the reference, deephar/models/spnet.py:283-296).  A new
    symbolic graph starts at every Input unless `graph` is given."""
    if shape is None and batch_shape is not None:
        shape = tuple(batch_shape)[1:]
    if shape is None:
        raise ValueError('Input: shape is required')
    shape = tuple(int(s) for s in shape)
    if graph is not None:
        g = graph
    else:
        g = Graph(name or 'model')
        g._counters = _SESSION_COUNTERS
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


class _SliceProbe(object):
    """What a Lambda body sees instead of a tensor: it may take a slice of the channel axis, nothing else."""

    def __init__(self, t):
        self._t = t

    def __getitem__(self, idx):
        idx = idx if isinstance(idx, tuple) else (idx,)
        full = slice(None, None, None)
        if len(idx) != len(self._t.shape) + 1 or any(i != full for i in idx[:-1]) or not isinstance(idx[-1], slice) \
                or idx[-1].step not in (None, 1):
            raise NotImplementedError('Lambda: only channel slices x[..., a:b] can be recorded')
        c = self._t.channels
        a, b, _ = idx[-1].indices(c)
        return L.channel_slice(self._t, a, b)

    def __getattr__(self, name):
        raise NotImplementedError('Lambda: only channel slices x[..., a:b] can be recorded (the body used .%s)' % name)


class Lambda(Layer):
    """keras.layers.Lambda, for the one body the reference's forward path needs outside its parameter-free head models:
    a slice of the channel axis (`Lambda(lambda x: x[:,:,:,:num_joints])`, reception.py:171-172).  Any other body is
    arbitrary backend arithmetic and is rejected."""

    def __init__(self, function, name=None, **kwargs):
        Layer.__init__(self, name, **{k: v for k, v in kwargs.items() if k not in ('output_shape', 'arguments', 'mask')})
        self.function = function

    def call(self, x):
        from .graph import Tensor
        try:
            out = self.function(_SliceProbe(x))
        except NotImplementedError:
            raise
        except Exception as e:                                  # K.* on the probe, arithmetic, ...
            raise NotImplementedError('Lambda: only channel slices x[..., a:b] can be recorded (%s)' % e)
        if not isinstance(out, Tensor):
            raise NotImplementedError('Lambda: only channel slices x[..., a:b] can be recorded')
        return out


class TimeDistributed(Layer):
    """keras.layers.TimeDistributed: frames already sit on the batch axis, so the wrapper only hands its name down
    (a wrapped nested Model keeps its own name: its weights live under that scope)."""

    def __init__(self, layer, name=None, **kwargs):
        Layer.__init__(self, name, **kwargs)
        self.layer = layer
        if name is not None and isinstance(layer, Layer):
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


# ---------------------------------------------------------------------------------------------------------------------
# The reference's parameter-free head models (deephar/models/blocks.py:217-343).  In Keras they are small sub-models of
# Lambda / pooling / frozen Dense and SeparableConv2D layers; here each is ONE recordable object, and the combination the
# builders make of them (reception.py:167-190) is rewritten into the single fused graph op the kernels implement.
# ---------------------------------------------------------------------------------------------------------------------
class _Head(object):
    def __init__(self, kind, name=None, **attrs):
        self.kind, self.name, self.attrs = kind, name, attrs
        self.trainable = False

    def __call__(self, x):
        xs = list(x) if isinstance(x, (list, tuple)) else [x]
        g = xs[0].g
        if self.kind in ('sam2d', 'jprob'):
            h, w, c = xs[0].shape
            shape = (1, c, 2 if self.kind == 'sam2d' else 1)
        elif self.kind == 'agg':
            shape = (1, self.attrs['num_joints'], 2)
        else:
            raise NotImplementedError('head model %r is not recordable yet' % self.kind)
        return g.op('head_' + self.kind, xs, shape, dict(self.attrs, name=self.name))


def build_softargmax_2d(input_shape, rho=0., name=None):
    """blocks.py:306-325: channel softmax + the two linear interpolations = soft-argmax of every map -> (C, 2)."""
    _no(rho > 0, 'kl_divergence_regularizer (rho > 0)')
    return _Head('sam2d', name)


def build_joints_probability(input_shape, name=None, verbose=0):
    """blocks.py:328-343: 4 * AveragePooling2D((2,2), strides 1) -> GlobalMaxPooling2D on the raw maps -> (C, 1)."""
    return _Head('jprob', name)


def build_context_aggregation(num_joints, num_context, alpha, num_frames=1, name=None):
    """blocks.py:217-285: pose = alpha * ys + (1 - alpha) * sum(pc * yc) / sum(pc) over each joint's context maps."""
    _no(num_frames != 1, 'build_context_aggregation(num_frames > 1)')
    return _Head('agg', name, num_joints=int(num_joints), num_context=int(num_context), alpha=float(alpha))


def build_softargmax_1d(input_shape, name=None):
    """blocks.py:288-303 (zSAM of the 3-D head).  Built by reception.build for every model, only CALLED for dim = 3."""
    return _Head('sam1d', name)


def _fuse_heads(g, outputs):
    """Replay the recorded graph into a new one, replacing the head-model calls by the fused ops of the layer graph:
       agg([sam2d(h[..., :nj]), sam2d(h[..., nj:]), jprob(h[..., nj:])]) + jprob(h[..., :nj])  ->  pose_regression_2d_context(h)
       sam2d(h) + jprob(h)                                                                  ->  pose_regression_2d(h)
    Returns (new graph, new outputs); a head call left over after the rewrite is an error."""
    if not any(nd.op.startswith('head_') for nd in g.nodes):
        return g, outputs

    def sliced(t):
        nd = t.node
        if nd is not None and nd.op == 'slice':
            return nd.inputs[0], nd.attrs['c0'], nd.attrs['c1']
        return t, 0, t.channels

    plans = {}          # node id -> ('ctx' | 'plain', ...) for the node at whose position the fused op is emitted
    alias = {}          # head node id -> (anchor node id, output index of the fused op)
    jprobs = [nd for nd in g.nodes if nd.op == 'head_jprob']
    used = set()
    for nd in g.nodes:
        if nd.op != 'head_agg':
            continue
        ys, yc, pc = nd.inputs
        if not (ys.node.op == 'head_sam2d' and yc.node.op == 'head_sam2d' and pc.node.op == 'head_jprob'):
            raise NotImplementedError('context aggregation of tensors that are not soft-argmax / probability heads')
        h, a0, a1 = sliced(ys.node.inputs[0])
        h2, b0, b1 = sliced(yc.node.inputs[0])
        h3, c0, c1 = sliced(pc.node.inputs[0])
        nj, nc = nd.attrs['num_joints'], nd.attrs['num_context']
        if not (h is h2 is h3 and (a0, a1) == (0, nj) and (b0, b1) == (c0, c1) == (nj, h.channels)
                and h.channels == nj * (nc + 1)):
            raise NotImplementedError('context aggregation over an unexpected split of the heat-maps')
        vis = [j for j in jprobs if j.id not in used and sliced(j.inputs[0])[0] is h and sliced(j.inputs[0])[1:] == (0, nj)]
        if len(vis) != 1:
            raise NotImplementedError('context head without its joint-probability on the specialised maps')
        for n_ in (ys.node, yc.node, pc.node, vis[0]):
            used.add(n_.id)
        plans[nd.id] = ('ctx', h, nd.attrs)
        alias[vis[0].id] = (nd.id, 1)
    for nd in g.nodes:
        if nd.op == 'head_sam2d' and nd.id not in used:
            t = nd.inputs[0]
            vis = [j for j in jprobs if j.id not in used and j.inputs[0] is t]
            if len(vis) != 1:
                raise NotImplementedError('soft-argmax head without its joint-probability model')
            used.update((nd.id, vis[0].id))
            plans[nd.id] = ('plain', t, nd.attrs)
            alias[vis[0].id] = (nd.id, 1)
    left = [nd for nd in g.nodes if nd.op.startswith('head_') and nd.id not in used and nd.id not in plans]
    if left:
        raise NotImplementedError('head model call(s) outside a recordable pattern: %s' % left)

    new = Graph(g.name)
    new.frames_per_clip = g.frames_per_clip
    new.weight_specs = list(g.weight_specs)
    new._weight_names = set(g._weight_names)
    new._counters = g._counters
    tmap, fused = {}, {}
    for nd in g.nodes:
        if nd.op == 'input':
            tmap[nd.outs[0].id] = new.input(nd.outs[0].shape, nd.outs[0].kind)
            continue
        if nd.id in plans:
            kind, h, attrs = plans[nd.id]
            if kind == 'ctx':
                outs = new.op('pose_regression_2d_context', [tmap[h.id]],
                              [(1, attrs['num_joints'], 2), (1, attrs['num_joints'], 1)],
                              {'num_joints': attrs['num_joints'], 'num_context': attrs['num_context'],
                               'alpha': attrs['alpha']})
            else:
                c = h.channels
                outs = new.op('pose_regression_2d', [tmap[h.id]], [(1, c, 2), (1, c, 1)], {})
            fused[nd.id] = outs
            tmap[nd.outs[0].id] = outs[0]
            continue
        if nd.id in alias:
            anchor, idx = alias[nd.id]
            if anchor not in fused:
                raise NotImplementedError('joint-probability head recorded before the soft-argmax head it belongs to')
            tmap[nd.outs[0].id] = fused[anchor][idx]
            continue
        if nd.op.startswith('head_'):
            continue                                    # absorbed operands (ps, pc, vc)
        attrs = {k: (dict(v) if isinstance(v, dict) else v) for k, v in nd.attrs.items()}
        outs = new.op(nd.op, [tmap[t.id] for t in nd.inputs], [o.shape for o in nd.outs], attrs, kind=nd.outs[0].kind)
        outs = outs if isinstance(outs, tuple) else (outs,)
        for o_src, o_new in zip(nd.outs, outs):
            tmap[o_src.id] = o_new
    return new, [tmap[t.id] for t in outputs]


def _weight_names(attrs, known):
    """(path in attrs, weight name) for every attribute value that names a weight of the recorded graph."""
    for key, val in attrs.items():
        if isinstance(val, str) and val in known:
            yield (key, None), val
        elif isinstance(val, dict):
            for k2, v2 in val.items():
                if isinstance(v2, str) and v2 in known:
                    yield (key, k2), v2


class Model(object):
    """keras.models.Model(inputs, outputs, name).

    * used as a MODEL (`predict`, `load_weights`, `outputs`, `weight_specs`, ...): the recorded graph is compiled for
      B200 on first use (layers that reach no output are dropped, as Keras drops them) and every attribute of
      deephar_b200.model.Model is available on this object;
    * used as a LAYER -- `Stem = Model(inp, x, name='Stem'); y = Stem(frames)`, the way the reference wraps its blocks
      (deephar/models/reception.py:96-98, 128-131) -- its layers are re-recorded into the caller's graph under the
      scope `name`, which is how Keras names the weights of a nested model in a checkpoint ("Stem/conv2d_1/kernel").
      A nested model can be applied once (no weight sharing on the reference's forward path).
    """

    def __init__(self, inputs=None, outputs=None, name=None):
        self._inputs = list(inputs) if isinstance(inputs, (list, tuple)) else [inputs]
        self._outputs = list(outputs) if isinstance(outputs, (list, tuple)) else [outputs]
        self._single = not isinstance(outputs, (list, tuple))
        if len(self._inputs) != 1:
            raise NotImplementedError('Model: exactly one Input (the frame / clip tensor) is supported')
        g = self._inputs[0].g
        for t in self._outputs:
            if t.g is not g:
                raise ValueError('Model: an output does not descend from the given Input')
        if g.inputs != self._inputs:
            raise ValueError('Model: `inputs` must be the Input the graph was started from')
        self._graph = g
        self._name = name or g.auto_name('model')
        self._impl = None
        self._applied = False

    # ---- as a model ---------------------------------------------------------------------------------------------
    def _compiled(self):
        if self._impl is None:
            if self._applied:
                raise NotImplementedError('Model %r was applied as a layer of another model; compile that one' % self._name)
            g, outs = _fuse_heads(self._graph, self._outputs)
            g.outputs = outs
            g.name = self._name
            self._impl = _Model(g, name=self._name)
        return self._impl

    def __getattr__(self, attr):
        # only reached for names this wrapper does not define: everything else is the compiled model's
        if attr.startswith('__') or attr in ('_impl', '_applied', '_graph', '_inputs', '_outputs', '_single', '_name'):
            raise AttributeError(attr)
        return getattr(self._compiled(), attr)

    # ---- as a layer ---------------------------------------------------------------------------------------------
    def __call__(self, x):
        if self._impl is not None or self._applied:
            raise NotImplementedError('Model %r: a nested model can be applied once, before it is used as a model'
                                      % self._name)
        sub, tgt = self._graph, x.g
        if tgt is sub:
            raise ValueError('Model %r applied to a tensor of its own graph' % self._name)
        src_in = self._inputs[0]
        if tuple(x.shape) != tuple(src_in.shape):
            raise ValueError('Model %r expects input shape %s, got %s' % (self._name, src_in.shape, x.shape))
        self._applied = True
        known = dict(sub.weight_specs)
        prefix = tgt.qualify(self._name)
        for wname, shape in sub.weight_specs:                        # creation order is kept
            layer, leaf = wname.rsplit('/', 1)
            tgt.add_weight(prefix + '/' + layer, leaf, shape)
        new = {src_in.id: x}
        for nd in sub.nodes:
            if nd.op == 'input':
                continue
            attrs = {}
            for k, v in nd.attrs.items():
                attrs[k] = dict(v) if isinstance(v, dict) else v
            if isinstance(attrs.get('name'), str) and nd.attrs.get('name') is not None and any(True for _ in _weight_names(nd.attrs, known)):
                attrs['name'] = prefix + '/' + attrs['name']
            for (k1, k2), w in _weight_names(nd.attrs, known):
                if k2 is None:
                    attrs[k1] = prefix + '/' + w
                else:
                    attrs[k1][k2] = prefix + '/' + w
            outs = tgt.op(nd.op, [new[t.id] for t in nd.inputs], [o.shape for o in nd.outs], attrs,
                          kind=nd.outs[0].kind if nd.outs[0].kind != src_in.kind else x.kind)
            outs = outs if isinstance(outs, tuple) else (outs,)
            for o_src, o_new in zip(nd.outs, outs):
                new[o_src.id] = o_new
        res = [new[t.id] for t in self._outputs]
        return res[0] if self._single else res
