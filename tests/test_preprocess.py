"""evaluation input pipeline (crop -> Pillow bilinear resize -> flip -> normalize_channels).

CPU: the oracle (oracle/preprocess.py) against the REFERENCE's own pipeline output (tests/golden/ref_preprocess.npz,
made by tests/golden/make_preprocess_golden.py) and, where Pillow is importable, against Pillow itself; the product's
host-side tables / boxes / affine maps against the oracle and the golden afmat.
GPU: dh_crop_resize_norm_u8 bit-exact against the golden frames and the oracle on ragged batches."""
import os

import numpy as np
import pytest

from deephar_b200 import preprocess
from oracle import preprocess as opre

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, 'golden', 'ref_preprocess.npz'))
CASES = ['down', 'down_flip', 'up', 'rect']


def _case(name):
    cx, cy, win, rw, rh, hflip = GOLD[name + '_args']
    return GOLD[name + '_src'], (cx, cy), win, (int(rw), int(rh)), int(hflip), GOLD[name + '_frame'], GOLD[name + '_afmat']


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_frames(name):
    src, objpos, win, res, hflip, frame, _ = _case(name)
    box = np.array([objpos[0] - win / 2, objpos[1] - win / 2, objpos[0] + win / 2, objpos[1] + win / 2], dtype=int)
    got = opre.eval_frame(src, box, res, hflip=bool(hflip))
    assert got.dtype == np.float32 and got.shape == frame.shape
    assert np.array_equal(got, frame)                                  # bit-exact float32


def test_oracle_matches_pillow():
    Image = pytest.importorskip('PIL.Image')
    rng = np.random.default_rng(5)
    for (h, w), size in [((97, 131), (64, 64)), ((40, 30), (96, 80)), ((256, 256), (256, 256)), ((300, 17), (17, 64)),
                         ((33, 500), (256, 8)), ((2, 2), (7, 5))]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize(size, Image.BILINEAR))
        assert np.array_equal(opre.resize_bilinear(img, size), want), ((h, w), size)
    img = rng.integers(0, 256, (50, 60, 3), dtype=np.uint8)
    for box in [(-10, -5, 30, 40), (20, 10, 90, 70), (5, 5, 25, 45)]:
        assert np.array_equal(opre.crop(img, box), np.asarray(Image.fromarray(img).crop(box)))


def test_product_tables_match_oracle():
    for a, b in [(150, 64), (40, 96), (7, 7), (1, 5), (513, 256), (180, 48), (3, 200), (1000, 3)]:
        pb, pc = preprocess.resample_tables(a, b)
        ob, oc = opre.resample_coefficients(a, b)
        assert np.array_equal(pb, ob) and np.array_equal(pc, oc), (a, b)
        assert pb.dtype == np.int32 and pc.dtype == np.int32


@pytest.mark.parametrize('name', CASES)
def test_product_box_and_afmat_match_reference(name):
    _, objpos, win, res, hflip, _, afmat = _case(name)
    box = preprocess.crop_box(objpos, win)
    got = preprocess.affine_map(box, res, hflip == 1)
    assert np.allclose(got, afmat, rtol=0, atol=1e-12)


def test_plan_shares_tables_and_rejects_bad_input():
    pipe = preprocess.FramePipeline((64, 64))
    frames, bounds, coefs, boxes, afmat, max_ch = pipe.plan([(100, 120)] * 3 + [(80, 90)], [[50, 50]] * 4,
                                                            [40.0, 40.0, 60.0, 40.0], hflip=[0, 1, 0, 0])
    assert frames[0].kx_off == frames[1].kx_off == frames[3].kx_off != frames[2].kx_off
    assert frames[0].kx_off == frames[0].ky_off                       # square window: one table for both axes
    assert max_ch == 60 and boxes.shape == (4, 4) and afmat.shape == (4, 3, 3)
    assert frames[1].hflip == 1 and frames[3].data == 3 * 100 * 120 * 3
    with pytest.raises(ValueError):
        pipe.plan([(10, 10)], [[5.5, 5.5]], [0.2])                      # truncates to an empty box
    with pytest.raises(NotImplementedError):
        pipe([np.zeros((8, 8, 3), np.uint8)], [[4, 4]], 4.0, angle=10)
    with pytest.raises(ValueError):
        pipe([np.zeros((8, 8, 3), np.float32)], [[4, 4]], 4.0)


def test_mpii_window_and_fixed_dataconf():
    """data/mpii.py:99-105 with config.py:42-50's fixed configuration."""
    from deephar_b200.config import DataConfig, mpii_sp_dataconf, pennaction_dataconf
    d = mpii_sp_dataconf.get_fixed_config()
    assert d == {'angle': 0, 'scale': 1, 'transx': 0, 'transy': 0, 'hflip': 0, 'chpower': 1, 'geoocclusion': None, 'subspl': 1}
    assert mpii_sp_dataconf.input_shape == (256, 256, 3) and pennaction_dataconf.get_fixed_config()['subspl'] == 6
    pos, win = preprocess.mpii_windows([[594.0, 257.0], [300.0, 100.0]], [3.021, 1.5], d)
    s = 1.25 * np.array([3.021, 1.5])
    assert np.allclose(pos, [[594.0, 257.0 + 12 * s[0]], [300.0, 100.0 + 12 * s[1]]], atol=1e-12)
    assert np.allclose(win, 200 * s, atol=1e-12)
    d2 = DataConfig(crop_resolution=(128, 128), scales=[0.7, 1.3], fixed_scale=1.3, fixed_trans_x=5).get_fixed_config()
    pos2, win2 = preprocess.mpii_windows([[10.0, 20.0]], [1.0], d2)
    assert np.allclose(pos2, [[10.0 + 1.25 * 5, 20.0 + 15.0]]) and np.allclose(win2, [200 * 1.3 * 1.25])
    with pytest.raises(NotImplementedError):
        mpii_sp_dataconf.random_data_generator()
    with pytest.raises(TypeError):
        DataConfig(resolution=(1, 1))


@pytest.mark.gpu
def test_gpu_pipeline_matches_reference_golden(cuda):
    for name in CASES:
        src, objpos, win, res, hflip, frame, afmat = _case(name)
        pipe = preprocess.FramePipeline(res)
        out, a = pipe([src], [objpos], win, hflip=hflip)
        assert np.array_equal(out.cpu().numpy()[0], frame), name
        assert np.allclose(a[0], afmat, rtol=0, atol=1e-12)
        assert pipe.launches == 2


@pytest.mark.gpu
def test_gpu_pipeline_ragged_batch_matches_oracle(cuda):
    rng = np.random.default_rng(11)
    shapes = [(120, 160), (90, 70), (256, 256), (33, 200), (64, 64), (300, 180), (17, 19)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    objpos = [[w * rng.uniform(0.1, 0.9), h * rng.uniform(0.1, 0.9)] for h, w in shapes]
    wins = [[rng.uniform(10, 1.5 * w), rng.uniform(10, 1.5 * h)] for h, w in shapes]
    wins[4] = [64.0, 64.0]
    objpos[4] = [32.0, 32.0]                                           # identity resize
    hflip = [0, 1, 0, 1, 0, 1, 0]
    for res, power in [((64, 64), 1), ((48, 80), (1.0, 0.9, 1.2))]:
        pipe = preprocess.FramePipeline(res)
        out, afmat = pipe(imgs, objpos, wins, hflip=hflip, channel_power=power)
        out = out.cpu().numpy()
        for i, im in enumerate(imgs):
            box = preprocess.crop_box(objpos[i], wins[i])
            want = opre.eval_frame(im, box, res, hflip=bool(hflip[i]), channel_power=power if power == 1 else list(power))
            if power == 1:
                assert np.array_equal(out[i], want), (i, res)
            else:                                                      # powf vs numpy.power: float32 ulps
                assert np.abs(out[i] - want).max() <= 4e-6, (i, res)
    if True:
        pipe = preprocess.FramePipeline((64, 64))
        out, afmat = pipe([], np.zeros((0, 2)), np.zeros((0,)))
        assert tuple(out.shape) == (0, 64, 64, 3) and afmat.shape == (0, 3, 3)


@pytest.mark.gpu
def test_gpu_pipeline_feeds_the_network_input_size(cuda):
    """256x256 evaluation resolution from larger images, batch of 32: output is the NHWC tensor predict() takes."""
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(32)]
    objpos = rng.uniform(200, 300, (32, 2))
    wins = rng.uniform(250, 500, 32)
    pipe = preprocess.FramePipeline((256, 256))
    out, afmat = pipe(imgs, objpos, wins)
    assert tuple(out.shape) == (32, 256, 256, 3) and out.is_contiguous()
    for i in (0, 13, 31):
        want = opre.eval_frame(imgs[i], preprocess.crop_box(objpos[i], wins[i]), (256, 256))
        assert np.array_equal(out[i].cpu().numpy(), want)


def test_tables_randomised_against_oracle_and_pillow():
    """Random (in, out) size pairs: product tables == oracle tables; oracle resize == Pillow where importable; every
    weight row sums to 2^22 within the rounding of its taps (Pillow's normalisation)."""
    rng = np.random.default_rng(2024)
    pairs = [(int(a), int(b)) for a, b in zip(rng.integers(1, 700, 40), rng.integers(1, 400, 40))]
    for a, b in pairs:
        pb, pc = preprocess.resample_tables(a, b)
        ob, oc = opre.resample_coefficients(a, b)
        assert np.array_equal(pb, ob) and np.array_equal(pc, oc), (a, b)
        assert np.all(pb[:, 0] >= 0) and np.all(pb[:, 0] + pb[:, 1] <= a) and np.all(pb[:, 1] >= 1)
        assert np.all(np.abs(pc.sum(axis=1) - (1 << 22)) <= pc.shape[1])
    try:
        from PIL import Image
    except ImportError:
        return
    for (h, w), (ow, oh) in [((int(a), int(b)), (int(c), int(d))) for a, b, c, d in rng.integers(1, 90, (12, 4))]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert np.array_equal(opre.resize_bilinear(img, (ow, oh)), want), ((h, w), (ow, oh))


def test_clip_sampling_and_windows_match_the_reference_helpers():
    """tests/golden/ref_clip_windows.npz: outputs of the reference's own get_clip_frame_index (random_clip=False) and
    bbox_to_objposwin / objposwin_to_bbox combined as data/pennaction.py:118-134 does (make_clipwindow_golden.py)."""
    G = np.load(os.path.join(HERE, 'golden', 'ref_clip_windows.npz'))
    for (size, sub, nf), want in zip(G['cases'], G['frames']):
        got = preprocess.clip_frame_index(int(size), int(sub), int(nf))
        assert got == [int(v) for v in want[:nf]], (size, sub, nf)
        assert len(got) == nf and all(0 <= f < size for f in got)
    for row in G['windows']:
        w, h, scale, tx, ty = row[:5]
        bbox = None if np.isnan(row[5]) else row[5:9]
        objpos, win = preprocess.clip_window((w, h), {'scale': scale, 'transx': tx, 'transy': ty}, bbox)
        assert np.allclose(objpos, row[9:11], rtol=0, atol=1e-12) and np.allclose(win, row[11:13], rtol=0, atol=1e-12)
    with pytest.raises(ValueError):
        preprocess.clip_frame_index(100, 0, 16)


@pytest.mark.timeout(180)
@pytest.mark.filterwarnings('ignore:This process .* is multi-threaded, use of fork')
def test_decode_images_keeps_order_and_content(tmp_path):
    """the batch front of FramePipeline: pooled Pillow decode == one Image.open per file, in the order given"""
    from PIL import Image
    from deephar_b200 import preprocess
    rng = np.random.default_rng(4)
    paths, want = [], []
    for i, (h, w) in enumerate([(48, 64), (30, 30), (64, 40), (33, 77), (20, 21)] * 3):
        arr = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        fmt = 'png' if i % 2 else 'jpg'
        path = str(tmp_path / ('im%d.%s' % (i, fmt)))
        Image.fromarray(arr).save(path, quality=95)
        paths.append(path)
        want.append(np.asarray(Image.open(path)))
    grey = str(tmp_path / 'grey.png')
    Image.fromarray(rng.integers(0, 256, (16, 18), dtype=np.uint8)).save(grey)
    paths.append(grey)
    want.append(np.asarray(Image.open(grey).convert('RGB')))

    def same(got):
        assert len(got) == len(want)
        for g, w_ in zip(got, want):
            assert g.dtype == np.uint8 and g.shape == w_.shape and np.array_equal(g, w_)
    same(preprocess.decode_images(paths))
    with preprocess.ImageDecoder(workers=3) as dec:
        same(dec(paths))
        same(dec(paths[::-1])[::-1])                # the pool is kept between calls
        assert dec([]) == [] and len(dec(paths[:1])) == 1
    assert dec._pool is None
