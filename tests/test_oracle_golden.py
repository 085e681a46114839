"""Pins the oracle against the committed fixtures (tests/golden/make_golden.py) and
cross-checks the numpy restatement against the independent torch-CPU implementation."""
import os

import numpy as np

from oracle import ops_np, ops_torch, reception, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_linspace_grid_matches_reference_source():
    """fixture = output of the reference's own linspace_2d source text (utils/math.py:6-19)."""
    z = np.load(os.path.join(GOLD, 'linspace_2d.npz'))
    for key in z.files:
        r, c, d = [int(s[1:]) for s in key.split('_')]
        got = ops_np.linspace_2d(r, c, dim=d)
        assert got.dtype == np.float32 and got.shape == z[key].shape
        assert np.array_equal(got, z[key]), key


def _small(ops, **kw):
    from deephar_b200.weights import load_calibration
    tab = synth.SyntheticTable(1234, load_calibration('reception_j16_d2_c2_k5'))
    x = synth.synth_frames(1, 64, 64, seed=3)
    return reception.forward(ops, tab, x, 16, 2, num_context_per_joint=2, num_blocks=2, ksize=(5, 5), **kw)


def test_oracle_reception_regression_pin():
    z = np.load(os.path.join(GOLD, 'reception_small_oracle.npz'))
    outs = _small(ops_np)
    assert len(outs) == len(z.files)
    for i, o in enumerate(outs):
        assert np.allclose(o, z['arr_%d' % i], rtol=0, atol=1e-9)


def test_numpy_vs_torch_cross_check():
    a = _small(ops_np)
    b = _small(ops_torch)
    for x, y in zip(a, b):
        # fp64 numpy vs fp32 oneDNN: coordinates agree to ~1e-5
        assert np.abs(x[..., :2] - y[..., :2]).max() < 2e-4
        assert np.abs(x[..., 2] - y[..., 2]).max() < 1e-3 * max(1.0, np.abs(x[..., 2]).max())


def test_3d_numpy_vs_torch():
    from deephar_b200.weights import load_calibration
    tab = synth.SyntheticTable(1234, load_calibration('reception_j17_d3_cNone_k5'))
    x = synth.synth_frames(1, 64, 64, seed=4)
    kw = dict(num_blocks=1, ksize=(5, 5), concat_pose_confidence=False)
    a = reception.forward(ops_np, tab, x, 17, 3, **kw)
    b = reception.forward(ops_torch, tab, x, 17, 3, **kw)
    assert a[0].shape == (1, 17, 3) and a[1].shape == (1, 17, 1)
    for x_, y_ in zip(a, b):
        assert np.abs(x_ - y_).max() < 2e-4
