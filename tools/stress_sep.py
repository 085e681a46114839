"""Race hunt: one fused separable layer launched many times on the same inputs; every output is compared with the
first one (bitwise) and with the fp64 oracle.  usage: python tools/stress_sep.py N H W Cin Cout k reps [share_a]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from deephar_b200 import _ffi, tc  # noqa: E402
from gpu_util import Dev, conv_desc  # noqa: E402
from oracle import ops_np  # noqa: E402

n, h, w, cin, cout, k, reps = [int(a) for a in sys.argv[1:8]]
dev = Dev(torch)
if len(sys.argv) > 8:
    dev.lib.dh_set_option(dev.ctx.handle, b'share_a', int(sys.argv[8]))
rng = np.random.default_rng(0)
x = rng.standard_normal((n, h, w, cin))
r0 = rng.standard_normal((n, h, w, cout))
dw = rng.standard_normal((k, k, cin, 1)) / k
pw = rng.standard_normal((1, 1, cin, cout)) / np.sqrt(cin)
post = (rng.uniform(0.5, 1.5, cout), rng.standard_normal(cout) * 0.3)
ref = ops_np.separable_conv2d(np.maximum(x, 0), dw, pw, (1, 1), 'same') * post[0] + post[1] + r0
xd, rd, dwd, pwd = dev.put(x), dev.put(r0), dev.put(dw), dev.put(pw)
hi, lo, cp, kp = tc.pack_matrix(pw.reshape(cin, cout).astype(np.float32))
th = torch.from_numpy(hi.view(np.int16).copy()).cuda()
tl = torch.from_numpy(lo.view(np.int16).copy()).cuda()
pk = _ffi.dh_packed_w(th.data_ptr(), tl.data_ptr(), cp, kp)
d = conv_desc(dev, (k, k), pre_relu=True, post=post, res=[dev.view(rd)], precision=3)
xv = dev.view(xd)
first = None
bad = 0
for i in range(reps):
    out = dev.empty(n, h, w, cout)
    ov = dev.view(out)
    if i % 3 == 2:                       # perturb the timing: cold L2 every third launch
        junk = torch.empty(64 << 20, device='cuda').fill_(1.0)
        del junk
    rc = dev.lib.dh_sepconv2d_f32(dev.ctx.handle, C.byref(xv), dwd.data_ptr(), pwd.data_ptr(), C.byref(pk), C.byref(d),
                                  C.byref(ov), dev.stream())
    _ffi.check(rc, 'sepconv')
    torch.cuda.synchronize()
    o = out.cpu().numpy()
    if first is None:
        first = o
        print('path', dev.lib.dh_last_conv_path(dev.ctx.handle), 'err vs oracle %.2e' % (np.abs(o - ref).max() / np.abs(ref).max()))
    diff = np.abs(o.astype(np.float64) - ref) > 1e-3 * np.abs(ref).max()
    if diff.any() or not np.array_equal(o, first):
        bad += 1
        idx = np.argwhere(diff)
        rows = sorted(set((int(a), int(b), int(c)) for a, b, c, _ in idx))
        cols = sorted(set(int(e) for _, _, _, e in idx))
        print('launch %d: %d elements off the oracle, %d differ from launch 0; pixels %s%s; channels %s%s'
              % (i, int(diff.sum()), int((o != first).sum()), rows[:6], '...' if len(rows) > 6 else '', cols[:12],
                 '...' if len(cols) > 12 else ''))
print('%d of %d launches bad' % (bad, reps))
