"""End-to-end coordinate error of the GPU forward vs the fp64/fp32 oracle for the tensor-core
precision modes (1 = plain bf16 operands, 3 = bf16x3 split) and the CUDA-core fp32 path."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deephar_b200 import reception  # noqa: E402
from oracle import ops_torch, synth  # noqa: E402
from oracle import reception as oracle_reception  # noqa: E402

res, blocks, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kw = dict(num_joints=16, dim=2, num_context_per_joint=2, num_blocks=blocks, ksize=(5, 5), concat_pose_confidence=False)
x = synth.synth_frames(n, res, res, seed=31)
ref = None
for mode in ('f32', 'tc3', 'tc1'):
    m = reception.build((res, res, 3), **kw)
    m.use_tensor_cores = mode != 'f32'
    m.precision = 1 if mode == 'tc1' else 3
    m.init_synthetic_weights(1234)
    if ref is None:
        dbg = {}
        ref = oracle_reception.forward(ops_torch, m.get_weights(), x, debug=dbg, **kw)
        cond = dbg['ctx_cond']
    outs = m.predict(x, batch_size=n)
    line = []
    for b in range(blocks):
        ok = cond[b] <= 30
        e = np.abs(outs[2 * b] - ref[2 * b]).max(-1)
        ev = np.abs(outs[2 * b + 1] - ref[2 * b + 1])[..., 0] / np.maximum(np.abs(ref[2 * b + 1][..., 0]), 1)
        line.append('%.1e/%.1e' % (e[ok].max(), ev.max()))
    print(mode, 'pose/vis err per block:', ' '.join(line))
