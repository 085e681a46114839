"""Third-party pins of the Keras / TF primitive semantics the oracle restates.

keras 2.1.4 / tensorflow 1.6 cannot be installed here, so the oracle's primitives were, until this file, checked only
against closed forms and against each other.  Two independent code bases that ARE in the image restate pieces of the same
semantics and are used here as referees:

  * HuggingFace `transformers` ships TensorFlow-compatible padding for its models ported from TF checkpoints
    (`models/mobilenet_v1/modeling_mobilenet_v1.py::apply_tf_padding`, `models/mobilenet_v2/...::apply_tf_padding`):
    TF "SAME" with the odd pixel on the bottom / right.  Its formula is written differently from the oracle's
    (`in % stride` cases vs `ceil(in / s)`), so agreement over a sweep is a real check of `ops_np.same_pad`.
  * torch's own `conv2d` / `max_pool2d` / `avg_pool2d` / `batch_norm` / `interpolate(nearest)` / `softmax` do the
    arithmetic, with the HuggingFace padding in front: the numpy oracle's Conv2D, SeparableConv2D stages, MaxPooling2D
    ('same' pads with -inf), AveragePooling2D, BatchNormalization (eps 1e-3), UpSampling2D and soft-max must equal them.

What stays unpinned: that Keras 2.1.4 maps `padding='same'` to exactly this TF rule, its HWIO / (kh,kw,C,1) kernel layouts
and its BatchNormalization default epsilon -- documented API facts, not executable here.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ops_np

mobilenet_v1 = pytest.importorskip('transformers.models.mobilenet_v1.modeling_mobilenet_v1')
mobilenet_v2 = pytest.importorskip('transformers.models.mobilenet_v2.modeling_mobilenet_v2')


def _hf_pads(h, w, kh, kw, sh, sw):
    """(top, bottom, left, right) as HuggingFace's apply_tf_padding pads a (1, 1, h, w) tensor of ones"""
    conv = torch.nn.Conv2d(1, 1, (kh, kw), stride=(sh, sw), bias=False)
    y = mobilenet_v1.apply_tf_padding(torch.ones(1, 1, h, w), conv)
    rows = (y[0, 0].sum(dim=1) > 0).nonzero().flatten()
    cols = (y[0, 0].sum(dim=0) > 0).nonzero().flatten()
    return int(rows[0]), y.shape[2] - 1 - int(rows[-1]), int(cols[0]), y.shape[3] - 1 - int(cols[-1])


def test_same_padding_rule_matches_huggingface_tf_padding():
    """every (size, kernel, stride) of the forward path and a sweep around them: 3x3 s2 on 256 (reception.py:64),
    7x7 s2 on 256 (spnet.py:322), 3x3 s2 max-pool (reception.py:74), (2,2) pools with stride (ts,2) on (T, nj) maps ..."""
    sizes = list(range(1, 24)) + [32, 64, 128, 255, 256]
    for h in sizes:
        for k in (1, 2, 3, 5, 7):
            for s in (1, 2, 3):
                if k < s:
                    continue                        # never used; TF and the HF helper both clamp at 0 there anyway
                out, before, after = ops_np.same_pad(h, k, s)
                t, b, l, r = _hf_pads(h, h, k, k, s, s)
                assert (before, after) == (t, b) == (l, r), (h, k, s)
                assert out == -(-h // s) == (h + before + after - k) // s + 1
    # rectangular kernels / strides (action head: (3,1), (3,5), pool stride (2,2) on (16, nj))
    for (h, w, kh, kw, sh, sw) in [(16, 16, 3, 1, 1, 1), (16, 20, 3, 5, 1, 1), (16, 17, 2, 2, 2, 2), (8, 9, 2, 2, 1, 2)]:
        ho, wo, pt, pb, pl, pr = ops_np._out_and_pad(h, w, kh, kw, sh, sw, 'same')
        assert (pt, pb, pl, pr) == _hf_pads(h, w, kh, kw, sh, sw)


def test_mobilenet_v2_variant_agrees_too():
    conv = torch.nn.Conv2d(1, 1, 3, stride=2, bias=False)
    x = torch.ones(1, 1, 255, 256)
    assert torch.equal(mobilenet_v1.apply_tf_padding(x, conv), mobilenet_v2.apply_tf_padding(x, conv))


def _nchw(a):
    return torch.from_numpy(np.ascontiguousarray(a)).permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).numpy()


@pytest.mark.parametrize('h,w,k,s', [(16, 16, 3, 2), (15, 17, 3, 2), (16, 16, 7, 2), (9, 12, 5, 1), (8, 8, 1, 1), (10, 7, 3, 3)])
def test_conv2d_same_equals_torch_conv_behind_tf_padding(h, w, k, s):
    rng = np.random.default_rng(h * 100 + k)
    x = rng.normal(size=(2, h, w, 5))
    wt = rng.normal(size=(k, k, 5, 4))
    conv = torch.nn.Conv2d(5, 4, k, stride=s, bias=False)
    ref = F.conv2d(mobilenet_v1.apply_tf_padding(_nchw(x), conv), torch.from_numpy(wt).permute(3, 2, 0, 1), stride=s)
    assert np.abs(ops_np.conv2d(x, wt, (s, s), 'same') - _nhwc(ref)).max() < 1e-12
    # depthwise stage of SeparableConv2D: kernel (kh, kw, C, 1), one filter per channel, same padding rule and stride
    dw = rng.normal(size=(k, k, 5, 1))
    ref = F.conv2d(mobilenet_v1.apply_tf_padding(_nchw(x), conv), torch.from_numpy(dw).permute(2, 3, 0, 1), stride=s, groups=5)
    assert np.abs(ops_np.depthwise_conv2d(x, dw, (s, s), 'same') - _nhwc(ref)).max() < 1e-12
    pw = rng.normal(size=(1, 1, 5, 6))
    ref = F.conv2d(ref, torch.from_numpy(pw).permute(3, 2, 0, 1))
    assert np.abs(ops_np.separable_conv2d(x, dw, pw, (s, s), 'same') - _nhwc(ref)).max() < 1e-12


@pytest.mark.parametrize('h,w,pool,stride', [(16, 16, 3, 2), (15, 15, 3, 2), (16, 17, 2, 2), (9, 9, 2, 2)])
def test_maxpool_same_pads_with_minus_infinity(h, w, pool, stride):
    rng = np.random.default_rng(7)
    x = rng.normal(size=(2, h, w, 3)) - 5.0                       # all negative: zero padding would win every border window
    conv = torch.nn.Conv2d(1, 1, pool, stride=stride, bias=False)
    t, b, l, r = _hf_pads(h, w, pool, pool, stride, stride)
    padded = F.pad(_nchw(x), (l, r, t, b), value=float('-inf'))
    assert padded.shape == mobilenet_v1.apply_tf_padding(_nchw(x), conv).shape
    ref = F.max_pool2d(padded, pool, stride)
    got = ops_np.maxpool2d(x, (pool, pool), (stride, stride), 'same')
    assert np.array_equal(got, _nhwc(ref))
    assert np.array_equal(ops_np.maxpool2d(x[:, :h - h % 2, :w - w % 2], (2, 2)),
                          _nhwc(F.max_pool2d(_nchw(x[:, :h - h % 2, :w - w % 2]), 2)))      # Keras default: valid, stride = pool


def test_elementwise_and_resampling_primitives_equal_torch():
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 6, 7, 4))
    gamma, beta, mean = rng.uniform(0.8, 1.2, 4), rng.normal(size=4), rng.normal(size=4)
    var = rng.uniform(0.5, 1.5, 4)
    ref = F.batch_norm(_nchw(x), torch.from_numpy(mean), torch.from_numpy(var), torch.from_numpy(gamma), torch.from_numpy(beta),
                       training=False, eps=1e-3)
    assert np.abs(ops_np.batchnorm(x, gamma, beta, mean, var) - _nhwc(ref)).max() < 1e-12
    ref = F.batch_norm(_nchw(x), torch.from_numpy(mean), torch.from_numpy(var), None, torch.from_numpy(beta), training=False, eps=1e-3)
    assert np.abs(ops_np.batchnorm(x, None, beta, mean, var) - _nhwc(ref)).max() < 1e-12          # scale=False (layers.py)
    assert np.array_equal(ops_np.upsample2d(x), _nhwc(F.interpolate(_nchw(x), scale_factor=2, mode='nearest')))
    assert np.abs(ops_np.avgpool2d_2x2_s1_valid(x) - _nhwc(F.avg_pool2d(_nchw(x), 2, stride=1))).max() < 1e-15
    assert np.abs(ops_np.softmax(x) - torch.softmax(torch.from_numpy(x), dim=-1).numpy()).max() < 1e-15
    assert np.abs(ops_np.sigmoid(x) - torch.sigmoid(torch.from_numpy(x)).numpy()).max() < 1e-15
    assert np.array_equal(ops_np.global_maxpool2d(x), torch.from_numpy(x).amax(dim=(1, 2)).numpy())
    pads = ((1, 2), (0, 3))
    assert np.array_equal(ops_np.zeropad2d(x, pads), _nhwc(F.pad(_nchw(x), (0, 3, 1, 2))))
