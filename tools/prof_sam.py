"""Launch the soft-argmax head kernel on N frames of (32,32,48) maps (for ncu / timing)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deephar_b200 import _ffi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = _ffi.Context(0)
lib = _ffi.lib()
h = torch.randn(n, 32, 32, 48, device='cuda') * 3.0
pose = torch.empty(n, 16, 2, device='cuda')
vis = torch.empty(n, 16, 1, device='cuda')
hv = _ffi.dh_view(h.data_ptr(), n, 32, 32, 48, 48)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _ffi.check(lib.dh_softargmax2d_ctx_f32(ctx.handle, C.byref(hv), 16, 2, C.c_float(0.8), pose.data_ptr(), vis.data_ptr(), st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    _ffi.check(lib.dh_softargmax2d_ctx_f32(ctx.handle, C.byref(hv), 16, 2, C.c_float(0.8), pose.data_ptr(), vis.data_ptr(), st))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print('softargmax2d_ctx n=%d: %.1f us/launch, %.0f GB/s' % (n, ms * 1e3, n * 196800 / ms / 1e6))
