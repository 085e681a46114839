"""Writes tests/golden/tiny_keras_weights.h5 with the pure-Python HDF5 writer (deephar_b200/hdf5.py): the
first stem layers (conv2d_1..4 + their BatchNormalization, ~100 K parameters as float16) of a 1-block
ReceptionNet laid out as keras 2.1.4 `save_weights` does (one group
per top-level layer / nested sub-model, `layer_names` / `weight_names` attributes, '<layer>/<leaf>:0'
datasets), plus the constants the reference assigns itself (soft-argmax grids) that a loader must skip.

    python tests/golden/make_h5_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from deephar_b200 import hdf5, reception  # noqa: E402

KW = dict(num_joints=4, dim=2, num_context_per_joint=2, num_blocks=1, ksize=(3, 3), concat_pose_confidence=False)
SHAPE = (32, 32, 3)


def main():
    m = reception.build(SHAPE, **KW).init_synthetic_weights(7)
    path = os.path.join(HERE, 'tiny_keras_weights.h5')
    # float16 storage keeps the committed fixture small; values are exactly representable after the cast
    table = {k: v.astype(np.float16).astype(np.float32) for k, v in m.get_weights().items()}
    m.set_weights(table)
    m.save_weights(path)
    # append what a real checkpoint also holds: the frozen soft-argmax layers (layers.py:160-200)
    from deephar_b200 import keras_h5
    entries, attrs = keras_h5.read_entries(path)
    with hdf5.Writer(path) as w:
        layers = []
        for g, wn, arr in entries:
            if g not in layers:
                layers.append(g)
        layers = ['Stem', 'sSAM', 'cSAM']
        w.set_attr('/', 'layer_names', np.array([s.encode() for s in layers]))
        w.set_attr('/', 'backend', 'tensorflow')
        w.set_attr('/', 'keras_version', '2.1.4')
        for g in layers[:-2]:
            ws = [(wn, a) for gg, wn, a in entries if gg == g]
            if g == 'Stem':       # keep the committed file small: a partial checkpoint (by_name=True target)
                keep = tuple('%s_%d/' % (k, i) for k in ('conv2d', 'batch_normalization') for i in (1, 2, 3, 4))
                ws = [(wn, a) for wn, a in ws if wn.startswith(keep)]
            else:
                continue
            w.create_group(g)
            w.set_attr(g, 'weight_names', np.array([n.encode() for n, _ in ws]))
            for wn, a in ws:
                w.create_dataset(g + '/' + wn, a.astype(np.float16))
        for g, c in (('sSAM', 4), ('cSAM', 8)):
            w.create_group(g)
            names = ['%s_x/depthwise_kernel:0' % g, '%s_x/pointwise_kernel:0' % g]
            w.set_attr(g, 'weight_names', np.array([n.encode() for n in names]))
            w.create_dataset(g + '/' + names[0], np.zeros((8, 8, c, 1), np.float16))
            w.create_dataset(g + '/' + names[1], np.eye(c, dtype=np.float16).reshape(1, 1, c, c))
    print('wrote %s (%d bytes)' % (path, os.path.getsize(path)))


if __name__ == '__main__':
    main()
