"""Eager mini-engine: tensors carry values; every layer call is recorded (Node) so that a Model can be
re-executed on new inputs (sub-models called on other tensors, TimeDistributed, predict)."""
import re
from collections import defaultdict

import numpy as np

_UIDS = defaultdict(int)
CREATED_LAYERS = []          # every layer with weights, in build (= weight creation) order


def reset_uids():
    _UIDS.clear()
    del CREATED_LAYERS[:]


def get_uid(prefix=''):
    _UIDS[prefix] += 1
    return _UIDS[prefix]


def to_snake_case(name):
    intermediate = re.sub('(.)([A-Z][a-z0-9]+)', r'\1_\2', name)
    insecure = re.sub('([a-z])([A-Z])', r'\1_\2', intermediate).lower()
    if insecure[0] != '_':
        return insecure
    return 'private' + insecure


class KTensor(object):
    """A float64 numpy value with a (None, ...) static shape and the node that produced it."""
    __array_priority__ = 100

    def __init__(self, value, node=None, index=0):
        self.value = np.asarray(value, dtype=np.float64)
        self._node = node
        self._index = index
        self._keras_shape = (None,) + tuple(self.value.shape[1:])

    @property
    def shape(self):
        return self._keras_shape

    def get_shape(self):
        return self._keras_shape

    # arithmetic used inside Lambda bodies
    def _v(self, o):
        return o.value if isinstance(o, KTensor) else o

    def __add__(self, o): return KTensor(self.value + self._v(o))
    def __radd__(self, o): return KTensor(self._v(o) + self.value)
    def __sub__(self, o): return KTensor(self.value - self._v(o))
    def __rsub__(self, o): return KTensor(self._v(o) - self.value)
    def __mul__(self, o): return KTensor(self.value * self._v(o))
    def __rmul__(self, o): return KTensor(self._v(o) * self.value)
    def __truediv__(self, o): return KTensor(self.value / self._v(o))
    def __rtruediv__(self, o): return KTensor(self._v(o) / self.value)
    def __neg__(self): return KTensor(-self.value)
    def __pow__(self, o): return KTensor(self.value ** self._v(o))
    def __getitem__(self, idx): return KTensor(self.value[idx])


def _aslist(x):
    if isinstance(x, (list, tuple)):
        return list(x)
    return [x]


class Node(object):
    def __init__(self, layer, inputs, was_list):
        self.layer = layer
        self.inputs = inputs
        self.was_list = was_list
        self.outputs = None


class Layer(object):
    def __init__(self, name=None, trainable=True, input_shape=None, **kwargs):
        if not name:
            prefix = to_snake_case(self.__class__.__name__)
            name = prefix + '_' + str(get_uid(prefix))
        self.name = name
        self.trainable = trainable
        self.built = False
        self._weights = []        # dicts: name, value, trainable, fixed
        self._inbound_nodes = []

    # ---- weights ----
    def add_weight(self, name, shape, trainable=True, init=0.0):
        w = {'name': name, 'value': np.full(tuple(int(s) for s in shape), init, dtype=np.float64),
             'trainable': trainable, 'fixed': False}
        self._weights.append(w)
        return w

    @property
    def weights(self):
        return [w for w in self._weights if w['trainable']] + [w for w in self._weights if not w['trainable']]

    def get_weights(self):
        return [w['value'].copy() for w in self.weights]

    def set_weights(self, values):
        ws = self.weights
        assert len(ws) == len(values), (self.name, len(ws), len(values))
        for w, v in zip(ws, values):
            assert w['value'].shape == np.shape(v), (self.name, w['name'], w['value'].shape, np.shape(v))
            w['value'] = np.array(v, dtype=np.float64)
            w['fixed'] = True          # assigned by the reference's own code (non-trainable constant)

    def count_params(self):
        return int(sum(w['value'].size for w in self._weights))

    # ---- call protocol ----
    def build(self, input_shapes):
        pass

    def call(self, values):
        raise NotImplementedError(self.__class__.__name__)

    def _ensure_built(self, vals, was_list):
        if not self.built:
            shapes = [(None,) + tuple(v.shape[1:]) for v in vals]
            self.build(shapes if was_list else shapes[0])
            self.built = True
            if self._weights:
                CREATED_LAYERS.append(self)

    def compute(self, vals, was_list):
        self._ensure_built(vals, was_list)
        out = self.call(vals if was_list else vals[0])
        return out

    def __call__(self, inputs):
        was_list = isinstance(inputs, (list, tuple))
        ins = _aslist(inputs)
        for t in ins:
            assert isinstance(t, KTensor), 'layer %s called on %r' % (self.name, type(t))
        out = self.compute([t.value for t in ins], was_list)
        node = Node(self, ins, was_list)
        self._inbound_nodes.append(node)
        if isinstance(out, (list, tuple)):
            outs = [KTensor(o, node, i) for i, o in enumerate(out)]
            node.outputs = outs
            return outs
        t = KTensor(out, node, 0)
        node.outputs = [t]
        return t

    # shapes of the first call
    def get_input_shape_at(self, i):
        n = self._inbound_nodes[i]
        s = [t.shape for t in n.inputs]
        return s if n.was_list else s[0]

    def get_output_shape_at(self, i):
        n = self._inbound_nodes[i]
        s = [t.shape for t in n.outputs]
        return s if len(s) > 1 else s[0]

    @property
    def input_shape(self):
        return self.get_input_shape_at(0)

    @property
    def output_shape(self):
        return self.get_output_shape_at(0)


class InputLayer(Layer):
    def __init__(self, shape, name=None):
        if not name:
            name = 'input_' + str(get_uid('input'))
        super(InputLayer, self).__init__(name=name)
        self.shape = tuple(shape)


def Input(shape=None, batch_shape=None, name=None, dtype=None, tensor=None):
    if shape is None:
        shape = tuple(batch_shape[1:])
    layer = InputLayer(shape, name=name)
    node = Node(layer, [], False)
    t = KTensor(np.zeros((1,) + tuple(int(s) for s in shape)), node, 0)
    node.outputs = [t]
    layer._inbound_nodes.append(node)
    return t


class Model(Layer):
    """Functional container: re-executes the recorded nodes between its inputs and outputs."""

    def __init__(self, inputs=None, outputs=None, name=None):
        if not name:
            name = 'model_' + str(get_uid('model'))
        Layer.__init__(self, name=name)
        self._in_list = isinstance(inputs, (list, tuple))
        self._out_list = isinstance(outputs, (list, tuple))
        self.inputs = _aslist(inputs)
        self.outputs = _aslist(outputs)
        self.built = True
        # layers in creation order of their first node reachable from the outputs
        self.layers = []
        seen = set()

        def visit(t):
            node = t._node
            if node is None or id(node) in seen:
                return
            seen.add(id(node))
            if t in self.inputs:
                if node.layer not in self.layers:
                    self.layers.append(node.layer)
                return
            for i in node.inputs:
                visit(i)
            if node.layer not in self.layers:
                self.layers.append(node.layer)

        import sys
        sys.setrecursionlimit(100000)
        for o in self.outputs:
            visit(o)

    # container weights (Keras: trainable of all layers, then non-trainable of all layers)
    @property
    def weights(self):
        tw, nw = [], []
        for l in self.layers:
            for w in l.weights:
                (tw if w['trainable'] else nw).append(w)
        return tw + nw

    def get_layer(self, name=None, index=None):
        if index is not None:
            return self.layers[index]
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError('No such layer: ' + str(name))

    @property
    def input(self):
        return self.inputs if self._in_list else self.inputs[0]

    @property
    def output(self):
        return self.outputs if self._out_list else self.outputs[0]

    def _run(self, vals):
        env = {}
        for t, v in zip(self.inputs, vals):
            env[id(t)] = np.asarray(v, dtype=np.float64)

        def ev(t):
            if id(t) in env:
                return env[id(t)]
            node = t._node
            assert node is not None and not isinstance(node.layer, InputLayer), \
                'tensor not reachable from the model inputs (layer %s)' % (node.layer.name if node else '?')
            ivals = [ev(i) for i in node.inputs]
            out = node.layer.compute(ivals, node.was_list)
            outs = list(out) if isinstance(out, (list, tuple)) else [out]
            for ot, ov in zip(node.outputs, outs):
                env[id(ot)] = ov
            return env[id(t)]

        return [ev(o) for o in self.outputs]

    def call(self, values):
        vals = values if isinstance(values, (list, tuple)) else [values]
        outs = self._run(vals)
        return outs if self._out_list else outs[0]

    def compute(self, vals, was_list):
        outs = self._run(vals)
        return outs if self._out_list else outs[0]

    def predict(self, x, batch_size=None, verbose=0):
        vals = x if isinstance(x, (list, tuple)) else [x]
        outs = self._run(vals)
        return outs if self._out_list else outs[0]

    def compile(self, *a, **k):
        pass

    def summary(self, *a, **k):
        pass

    @property
    def input_shape(self):
        s = [t.shape for t in self.inputs]
        return s if self._in_list else s[0]

    @property
    def output_shape(self):
        s = [t.shape for t in self.outputs]
        return s if self._out_list else s[0]

    def get_input_shape_at(self, i):
        if self._inbound_nodes:
            return Layer.get_input_shape_at(self, i)
        return self.input_shape
