"""keras_compat: model code written in the Keras functional style (the style of deephar/models/*.py) records the same
graph, weight list and kernel plan as the function-style mirror of deephar/layers.py; unsupported Keras features are
rejected when the model is built."""
import numpy as np
import pytest

from deephar_b200 import keras_compat as K
from deephar_b200 import layers as L
from deephar_b200.graph import Graph
from deephar_b200.model import Model
from oracle import ops_torch


def _keras_style(inp):
    """A stem + separable residual unit + hourglass step, as one would write it for Keras."""
    x = K.Conv2D(32, (3, 3), strides=(2, 2), padding='same', use_bias=False, name='conv1')(inp)
    x = K.BatchNormalization(scale=False, name='bn1')(x)
    x = K.Activation('relu', name='act1')(x)
    a = K.SeparableConv2D(64, (5, 5), padding='same', use_bias=False, name='sep1')(K.Activation('relu')(x))
    a = K.BatchNormalization(name='bn2')(a)
    s = K.Conv2D(64, (1, 1), padding='same', use_bias=False, name='short')(x)
    x = K.add([s, a])
    low = K.MaxPooling2D((2, 2), strides=(2, 2), padding='same')(x)
    low = K.TimeDistributed(K.SeparableConv2D(64, (3, 3), padding='same', use_bias=False), name='td_low')(low)
    x = K.add([x, K.UpSampling2D((2, 2))(low)])
    h = K.Conv2D(8, (1, 1), use_bias=False, name='heat')(K.Activation('relu')(x))
    hm = K.channel_softmax_2d(h)
    return [K.softargmax2d(hm), K.keypoint_confidence(hm), K.concatenate([x, h])]


def _function_style(inp):
    x = L.conv2d(inp, 32, (3, 3), strides=(2, 2), name='conv1')
    x = L.relu(L.BatchNormalization(x, scale=False, name='bn1'), name='act1')
    a = L.BatchNormalization(L.sepconv2d(L.relu(x), 64, (5, 5), name='sep1'), name='bn2')
    s = L.conv2d(x, 64, (1, 1), name='short')
    x = L.add([s, a])
    low = L.sepconv2d(L.maxpooling2d(x), 64, (3, 3), name='td_low')
    x = L.add([x, L.UpSampling2D(low)])
    h = L.conv2d(L.relu(x), 8, (1, 1), padding='valid', name='heat')
    hm = L.channel_softmax_2d(h)
    return [L.softargmax2d(hm), L.keypoint_confidence(hm), L.concatenate([x, h])]


def _models():
    inp = K.Input(shape=(32, 32, 3))
    mk = K.Model(inputs=inp, outputs=_keras_style(inp), name='toy')
    g = Graph('toy')
    g.outputs = _function_style(g.input((32, 32, 3)))
    return mk, Model(g)


def test_keras_style_records_the_same_model():
    mk, mf = _models()
    assert mk.weight_specs == mf.weight_specs
    assert ('td_low/depthwise_kernel', (3, 3, 64, 1)) in mk.weight_specs
    assert [k.kind for k in mk.plan.kops] == [k.kind for k in mf.plan.kops]
    assert mk.output_shape == mf.output_shape == [(None, 8, 2), (None, 8, 1), (None, 16, 16, 72)]
    assert mk.name == 'toy' and mk.input_shape == (None, 32, 32, 3)


def test_clip_input_folds_time():
    inp = K.Input(shape=(4, 16, 16, 3))
    assert inp.g.frames_per_clip == 4 and inp.shape == (16, 16, 3)


def test_unsupported_keras_features_fail_at_build_time():
    inp = K.Input(shape=(16, 16, 3))
    with pytest.raises(NotImplementedError):
        K.Conv2D(8, (3, 3))                                   # Keras default use_bias=True
    with pytest.raises(NotImplementedError):
        K.Conv2D(8, (3, 3), use_bias=False, activation='relu')
    with pytest.raises(NotImplementedError):
        K.SeparableConv2D(8, (3, 3), use_bias=False, depth_multiplier=2)
    with pytest.raises(NotImplementedError):
        K.Activation('tanh')
    with pytest.raises(NotImplementedError):
        K.BatchNormalization(epsilon=1e-5)
    with pytest.raises(TypeError):
        K.UpSampling2D(interpolation='bilinear')
    shared = K.Conv2D(3, (1, 1), use_bias=False)
    y = shared(inp)
    with pytest.raises(NotImplementedError):
        shared(y)
    other = K.Input(shape=(16, 16, 3))
    with pytest.raises(ValueError):
        K.Model(inputs=other, outputs=[y])


@pytest.mark.gpu
def test_keras_style_model_runs_and_matches_oracle(cuda):
    mk, _ = _models()
    mk.init_synthetic_weights(7)
    x = np.random.default_rng(0).uniform(-1, 1, (3, 32, 32, 3)).astype(np.float32)
    outs = mk.predict(x)
    refs = _oracle_forward(mk.get_weights(), x)
    for o, r in zip(outs, refs):
        assert o.shape == r.shape and np.abs(o - r).max() <= 1e-4


def _oracle_forward(w, x):
    o = ops_torch
    w = {k: o.from_numpy(v) for k, v in w.items()}

    def bn(t, name, scale=True):
        return o.batchnorm(t, w[name + '/gamma'] if scale else None, w[name + '/beta'], w[name + '/moving_mean'],
                           w[name + '/moving_variance'])

    t = o.conv2d(o.from_numpy(x), w['conv1/kernel'], (2, 2), 'same')
    t = o.relu(bn(t, 'bn1', scale=False))
    a = bn(o.separable_conv2d(o.relu(t), w['sep1/depthwise_kernel'], w['sep1/pointwise_kernel']), 'bn2')
    t = o.conv2d(t, w['short/kernel']) + a
    low = o.maxpool2d(t, (2, 2), (2, 2), 'same')
    low = o.separable_conv2d(low, w['td_low/depthwise_kernel'], w['td_low/pointwise_kernel'])
    t = t + o.upsample2d(low)
    h = o.conv2d(o.relu(t), w['heat/kernel'], (1, 1), 'valid')
    hm = o.channel_softmax_2d(h)
    pose = o.to_numpy(o.softargmax2d(hm)).reshape(-1, 8, 2)
    vis = o.to_numpy(o.keypoint_confidence(hm)).reshape(-1, 8, 1)
    return [pose, vis, o.to_numpy(o.concat([t, h]))]
