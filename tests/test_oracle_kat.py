"""Closed-form known-answer tests for the oracle ops (SURVEY.md 8c list) -- run for both the
numpy fp64 restatement and the independent torch-CPU implementation."""
import numpy as np
import pytest

from oracle import ops_np, ops_torch

BACKENDS = [ops_np, ops_torch]


def _arr(ops, a):
    return ops.from_numpy(np.asarray(a, dtype=np.float64))


@pytest.mark.parametrize('ops', BACKENDS)
def test_softargmax_onehot_and_uniform(ops):
    H, W, C = 6, 9, 3
    x = np.full((1, H, W, C), -50.0)
    peaks = [(0, 0), (5, 8), (2, 3)]
    for c, (r, q) in enumerate(peaks):
        x[0, r, q, c] = 50.0
    p = ops.softargmax2d(ops.channel_softmax_2d(_arr(ops, x)))
    p = ops.to_numpy(p)
    for c, (r, q) in enumerate(peaks):
        assert np.allclose(p[0, c], [q / (W - 1), r / (H - 1)], atol=1e-6)
    u = ops.to_numpy(ops.softargmax2d(ops.channel_softmax_2d(_arr(ops, np.zeros((1, H, W, C))))))
    assert np.allclose(u, 0.5, atol=1e-6)


@pytest.mark.parametrize('ops', BACKENDS)
def test_softargmax_1d(ops):
    D, C = 16, 4
    x = np.full((1, D, C), -60.0)
    for c in range(C):
        x[0, 3 * c + 1, c] = 60.0
    z = ops.to_numpy(ops.lin_interpolation_1d(ops.channel_softmax_1d(_arr(ops, x))))
    for c in range(C):
        assert abs(z[0, c, 0] - (3 * c + 1 + 0.5) / D) < 1e-6
    u = ops.to_numpy(ops.lin_interpolation_1d(ops.channel_softmax_1d(_arr(ops, np.zeros((1, D, C))))))
    assert np.allclose(u, 0.5, atol=1e-6)


@pytest.mark.parametrize('ops', BACKENDS)
def test_keypoint_confidence(ops):
    R = 8
    onehot = np.zeros((1, R, R, 1))
    onehot[0, 3, 4, 0] = 1.0
    assert abs(ops.to_numpy(ops.keypoint_confidence(_arr(ops, onehot)))[0, 0, 0] - 1.0) < 1e-6
    uni = np.full((1, R, R, 1), 1.0 / R ** 2)
    assert abs(ops.to_numpy(ops.keypoint_confidence(_arr(ops, uni)))[0, 0, 0] - 4.0 / R ** 2) < 1e-7


@pytest.mark.parametrize('ops', BACKENDS)
def test_sepconv_identity_and_same_padding(ops):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 7, 6, 5))
    dw = np.zeros((5, 5, 5, 1))
    dw[2, 2, :, 0] = 1.0
    pw = np.eye(5).reshape(1, 1, 5, 5)
    y = ops.to_numpy(ops.separable_conv2d(_arr(ops, x), _arr(ops, dw), _arr(ops, pw)))
    assert np.allclose(y, x, atol=1e-6)
    # TF SAME, 3x3 stride 2 on a 4x4 ramp: pads only bottom/right
    ramp = np.arange(16, dtype=np.float64).reshape(1, 4, 4, 1)
    w = np.ones((3, 3, 1, 1))
    y = ops.to_numpy(ops.conv2d(_arr(ops, ramp), _arr(ops, w), (2, 2), 'same'))[0, :, :, 0]
    ref = np.array([[ramp[0, 0:3, 0:3, 0].sum(), ramp[0, 0:3, 2:4, 0].sum()],
                    [ramp[0, 2:4, 0:3, 0].sum(), ramp[0, 2:4, 2:4, 0].sum()]])
    assert np.allclose(y, ref)
    assert ops_np.same_pad(256, 3, 2) == (128, 0, 1)
    assert ops_np.same_pad(256, 7, 2) == (128, 2, 3)


@pytest.mark.parametrize('ops', BACKENDS)
def test_batchnorm_scale_false(ops):
    x = np.array([[[[1.0, -2.0]]]])
    mean, var, beta = np.array([0.5, 1.0]), np.array([4.0, 0.25]), np.array([0.1, -0.1])
    y = ops.to_numpy(ops.batchnorm(_arr(ops, x), None, _arr(ops, beta), _arr(ops, mean), _arr(ops, var)))
    assert np.allclose(y[0, 0, 0], (x[0, 0, 0] - mean) / np.sqrt(var + 1e-3) + beta, atol=1e-6)


@pytest.mark.parametrize('ops', BACKENDS)
def test_max_min_pooling_and_maxpool_same(ops):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((1, 5, 7, 3))
    y = ops.to_numpy(ops.max_min_pooling(_arr(ops, x)))
    mp = ops.to_numpy(ops.maxpool2d(_arr(ops, x), (2, 2), None, 'same'))
    mn = -ops.to_numpy(ops.maxpool2d(_arr(ops, -x), (2, 2), None, 'same'))
    assert y.shape == (1, 3, 4, 3)
    assert np.allclose(y, mp + mn, atol=1e-6)
    # -inf padding: a negative map keeps its (negative) max at the padded border
    neg = -np.ones((1, 3, 3, 1))
    assert np.allclose(ops.to_numpy(ops.maxpool2d(_arr(ops, neg), (3, 3), (2, 2), 'same')), -1.0)


@pytest.mark.parametrize('ops', BACKENDS)
def test_kronecker_onehot(ops):
    rng = np.random.default_rng(2)
    z = rng.standard_normal((1, 2, 4, 4, 6))
    p = np.zeros((1, 2, 4, 4, 3))
    p[0, :, 1, 2, 0] = 1
    p[0, :, 3, 3, 1] = 1
    p[0, :, 0, 0, 2] = 1
    k = ops.to_numpy(ops.kronecker_prod(_arr(ops, p), _arr(ops, z)))
    assert np.allclose(k[0, :, 0], z[0, :, 1, 2], atol=1e-6)
    assert np.allclose(k[0, :, 1], z[0, :, 3, 3], atol=1e-6)
    assert np.allclose(k[0, :, 2], z[0, :, 0, 0], atol=1e-6)


def test_context_aggregation_equal_probabilities():
    from oracle import reception
    rng = np.random.default_rng(3)
    nj, nc = 4, 2
    ys = rng.uniform(size=(2, nj, 2))
    yc = rng.uniform(size=(2, nj * nc, 2))
    pc = np.full((2, nj * nc, 1), 0.37)
    y = reception.context_aggregation_model(ops_np, ys, yc, pc, nj, nc, 0.8)
    ref = 0.8 * ys + 0.2 * yc.reshape(2, nj, nc, 2).mean(axis=2)
    assert np.allclose(y, ref)


def test_3d_channel_order_is_depth_major():
    from oracle import reception
    nj, D = 3, 16
    h = np.full((1, 4, 4, D * nj), -40.0)
    for j, d in enumerate([0, 7, 15]):
        h[0, :, :, d * nj + j] = 40.0
    pose, vis, _ = reception.pose_regression_3d(ops_np, h, nj, D)
    for j, d in enumerate([0, 7, 15]):
        assert abs(pose[0, j, 2] - (d + 0.5) / D) < 1e-6
