"""A stand-in CUDA device for HOST-LOGIC tests -- TEST INFRASTRUCTURE (never imported by the product).

    python tests/fake_cuda.py bench.py --gpus 1 --steps 1 ...          (or under torch.distributed.run)

`install()` makes the host side of deephar_b200 and bench.py runnable on a machine without a GPU so that their control
flow can be exercised: `torch.cuda` reports one device whose "device memory" is host memory (`device='cuda'` allocations,
`.cuda()`, `.pin_memory()` become host tensors), streams / events / graphs are inert objects (an event pair reports wall
time), `init_process_group('nccl')` becomes gloo, and the C library keeps its pure host entry points (`dh_tc_k_pad`,
`dh_version`, ...) while EVERY DEVICE ENTRY POINT IS A NO-OP RETURNING 0.  No arithmetic is done: outputs are whatever the
buffers held.  What this covers is everything around the kernels -- batching, shard arithmetic, buffer binding, descriptor
construction (ctypes argument types are still checked by the stubs' signatures), copy pipelining, the collective plumbing
and the JSON line of bench.py -- at any world size.  Numbers printed under it mean nothing and are labelled as such.

`install(arithmetic=True)` goes one step further for the entry points of the model forward (convolutions, pooling,
upsampling, adds, soft-argmax heads, kronecker product, action-head ops): the stubs decode their raw arguments -- `dh_view`
pointers with channel offsets and leading dimensions, `dh_conv_desc` with its scale / shift / residual / pooled-output
fields, weight pointers into the model's flat arena -- exactly as the CUDA side does, and compute the op with the oracle's
numpy primitives in float64, rounding to the fp32 buffers.  `Model.predict` then produces real numbers on the CPU through
the product's OWN `_bind` / `_issue` / copy pipeline, which tests/test_host_path.py compares with the oracle.  Still test
infrastructure: the product never imports this file and has no CPU path.
"""
import contextlib
import ctypes as C
import os
import runpy
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HOST_ENTRY_POINTS = ('dh_tc_k_pad', 'dh_tc_cout_pad', 'dh_version', 'dh_last_error')
CAPTURING = []          # the stand-in CUDA graph being captured, if any


def _is_cuda(dev):
    return dev is not None and str(dev).startswith('cuda')


# ---- optional arithmetic behind the forward-path entry points (see the module docstring) -----------------------------------
def _f32(ptr, count):
    import numpy as np
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(int(count),))


def _view(vp):
    """dh_view* -> writable float32 numpy view (n, h, w, c) with the view's leading dimension"""
    import numpy as np
    v = vp.contents
    n, h, w, c, ld = int(v.n), int(v.h), int(v.w), int(v.c), int(v.ld)
    flat = _f32(v.p, (n * h * w - 1) * ld + c)
    return np.lib.stride_tricks.as_strided(flat, shape=(n, h, w, c), strides=(h * w * ld * 4, w * ld * 4, ld * 4, 4))


def _struct_view(v):
    return _view(C.pointer(v))


def _conv(separable):
    def op(ctx, x, *rest):
        import numpy as np
        from oracle import ops_np as O
        if separable:
            w_dw, w_pw, packed, d, out, stream = rest
        else:
            w, packed, d, out, stream = rest
        d = d.contents
        a = _view(x).astype(np.float64)
        o = _view(out)
        cin, cout = a.shape[-1], o.shape[-1]
        if d.pre_scale:
            a = a * _f32(d.pre_scale, cin) + _f32(d.pre_shift, cin)
        if d.pre_relu:
            a = np.maximum(a, 0.0)
        strides, padding = (int(d.sh), int(d.sw)), 'same' if d.pad_same else 'valid'
        kh, kw = int(d.kh), int(d.kw)
        if separable:
            y = O.separable_conv2d(a, _f32(w_dw, kh * kw * cin).reshape(kh, kw, cin, 1).astype(np.float64),
                                   _f32(w_pw, cin * cout).reshape(1, 1, cin, cout).astype(np.float64), strides, padding)
        else:
            y = O.conv2d(a, _f32(w, kh * kw * cin * cout).reshape(kh, kw, cin, cout).astype(np.float64), strides, padding)
        if d.post_scale:
            y = y * _f32(d.post_scale, cout) + _f32(d.post_shift, cout)
        if d.post_relu:
            y = np.maximum(y, 0.0)
        for i in range(int(d.n_res)):
            r = _struct_view(d.res[i]).astype(np.float64)
            if (int(d.res_up2x) >> i) & 1:
                r = O.upsample2d(r)
            y = y + r
        assert y.shape == o.shape, (y.shape, o.shape)
        o[...] = y
        if d.pool_out.p:
            _struct_view(d.pool_out)[...] = O.maxpool2d(y, (2, 2))
    return op


def _maxpool(ctx, x, kh, kw, sh, sw, pad_same, out, stream):
    from oracle import ops_np as O
    _view(out)[...] = O.maxpool2d(_view(x).astype('float64'), (kh, kw), (sh, sw), 'same' if pad_same else 'valid')


def _upsample_add(ctx, a, b, out, stream):
    from oracle import ops_np as O
    y = O.upsample2d(_view(b).astype('float64'))
    if a and a.contents.p:
        y = y + _view(a)
    _view(out)[...] = y


def _add_n(ctx, ins, n_in, scale, shift, relu, out, stream):
    import numpy as np
    o = _view(out)
    y = sum(_view(C.pointer(ins[i])).astype(np.float64) for i in range(n_in))
    if scale:
        y = y * _f32(scale, o.shape[-1]) + _f32(shift, o.shape[-1])
    if relu:
        y = np.maximum(y, 0.0)
    o[...] = y


def _dense(ptr, *shape):
    import numpy as np
    return _f32(ptr, int(np.prod(shape))).reshape(shape)


def _softargmax2d(ctx, h, d, alpha, conf_on_prob, out_pose, out_conf, prob_out, stream):
    import numpy as np
    from oracle import ops_np as O
    x = _view(h).astype(np.float64)
    n, _, _, c = x.shape
    p = O.channel_softmax_2d(x, float(alpha))
    pose = O.softargmax2d(p)
    if d and d.contents.p:
        pose = np.concatenate([pose, np.sum(O.sigmoid(_view(d).astype(np.float64)) * p, axis=(1, 2))[..., None]], axis=-1)
    _dense(out_pose, n, c, pose.shape[-1])[...] = pose
    _dense(out_conf, n, c, 1)[...] = O.keypoint_confidence(p if conf_on_prob else x)
    if prob_out and prob_out.contents.p:
        _view(prob_out)[...] = p


def _softargmax2d_ctx(ctx, h, nj, n_ctx, alpha_mix, out_pose, out_vis, stream):
    import numpy as np
    from oracle import ops_np as O
    x = _view(h).astype(np.float64)
    n = x.shape[0]
    hs, hc = x[..., :nj], x[..., nj:]
    ys, yc = O.softargmax2d(O.channel_softmax_2d(hs)), O.softargmax2d(O.channel_softmax_2d(hc))
    pc = O.keypoint_confidence(hc)
    grp = lambda v: v.reshape(n, nj, n_ctx, -1).sum(axis=2)     # noqa: E731
    with np.errstate(divide='ignore', invalid='ignore'):
        _dense(out_pose, n, nj, 2)[...] = alpha_mix * ys + (1 - alpha_mix) * grp(yc * pc) / grp(pc)
    _dense(out_vis, n, nj, 1)[...] = O.keypoint_confidence(hs)


def _softargmax3d_ex(ctx, h, nj, depth, vis_scale, out_pose, out_vis, prob_out, stream):
    import numpy as np
    from oracle import ops_np as O
    x = _view(h).astype(np.float64)
    n, hh, ww, _ = x.shape
    h5 = x.reshape(n, hh, ww, depth, nj)
    hxy, hz = h5.mean(axis=3), h5.mean(axis=(1, 2))
    _dense(out_pose, n, nj, 3)[...] = np.concatenate([O.softargmax2d(O.channel_softmax_2d(hxy)),
                                                      O.lin_interpolation_1d(O.channel_softmax_1d(hz))], axis=-1)
    _dense(out_vis, n, nj, 1)[...] = O.sigmoid(vis_scale * (hxy.max(axis=(1, 2)) + hz.max(axis=1)))[..., None]
    if prob_out and prob_out.contents.p:
        _view(prob_out)[...] = O.channel_softmax_2d(hxy)


def _softargmax3d(ctx, h, nj, depth, out_pose, out_vis, stream):
    _softargmax3d_ex(ctx, h, nj, depth, 1.0, out_pose, out_vis, None, stream)


def _kron(ctx, p, z, out, stream):
    import numpy as np
    a, b = _view(p).astype(np.float64), _view(z).astype(np.float64)
    _dense(out, a.shape[0], a.shape[-1], b.shape[-1])[...] = np.einsum('nhwj,nhwf->njf', a, b)


def _zeropad(ctx, x, top, left, out, stream):
    a, o = _view(x), _view(out)
    o[...] = 0.0
    o[:, top:top + a.shape[1], left:left + a.shape[2], :] = a


def _maxmin_pool(ctx, x, out, stream):
    from oracle import ops_np as O
    _view(out)[...] = O.max_min_pooling(_view(x).astype('float64'), (2, 2), 'same')


def _global_maxmin_softmax(ctx, x, out, stream):
    from oracle import ops_np as O
    a = _view(x).astype('float64')
    _dense(out, a.shape[0], a.shape[-1])[...] = O.softmax(O.global_max_min_pooling(a))


def _mask_mul(ctx, p, c, rows, dim, out, stream):
    _dense(out, rows, dim)[...] = _dense(p, rows, dim) * _dense(c, rows, 1)


def _pose_eval(ctx, pred, ldp, afmat, per_sample, inverse, y_true, head, refp, n, nj, out_pose, hits, valid, dsum, stream):
    """csrc/postprocess.cu: (inverse) affine of the poses, per-joint hit / valid counters and distance sums accumulated"""
    import numpy as np
    P = _dense(pred, n, nj, ldp).astype(np.float64)
    A = _dense(afmat, n if per_sample else 1, 3, 3).astype(np.float64)
    M = np.linalg.inv(A) if inverse else A
    M = np.broadcast_to(M, (n, 3, 3))
    t = np.einsum('nij,nkj->nki', M[:, :2, :2], P[:, :, :2]) + M[:, None, :2, 2]
    _dense(out_pose, n, nj, 2)[...] = t
    if y_true:
        g = _dense(y_true, n, nj, 2).astype(np.float64)
        ok = (g[..., 0] > -1e6) & (g[..., 1] > -1e6)
        d = np.sqrt(((g - t) ** 2).sum(axis=-1))
        np.ctypeslib.as_array(C.cast(valid, C.POINTER(C.c_int32)), shape=(nj,))[...] += ok.sum(axis=0).astype(np.int32)
        np.ctypeslib.as_array(C.cast(dsum, C.POINTER(C.c_double)), shape=(nj,))[...] += np.where(ok, d, 0.0).sum(axis=0)
        if head:
            d = d / _dense(head, n, 1).astype(np.float64)
        hit = ok & (d <= np.float64(np.float32(refp)))
        np.ctypeslib.as_array(C.cast(hits, C.POINTER(C.c_int32)), shape=(nj,))[...] += hit.sum(axis=0).astype(np.int32)


def _crop_resize_norm(ctx, table, n, max_ch, bounds, coefs, rh, rw, power, tmp, tmp_stride, out, stream):
    """csrc/preprocess.cu: per frame crop window (outside the image = 0) -> Pillow's bilinear resize -> flip ->
    normalize_channels; the resampling tables the host sent are not read here: the oracle derives its own"""
    import numpy as np
    from deephar_b200._ffi import dh_frame_src
    from oracle import preprocess as OP
    frames = C.cast(table, C.POINTER(dh_frame_src))
    o = _dense(out, n, rh, rw, 3)
    pw = 1 if not power else tuple(float(v) for v in _f32(power, 3))
    for i in range(n):
        f = frames[i]
        img = np.ctypeslib.as_array(C.cast(f.data, C.POINTER(C.c_uint8)), shape=(int(f.h), int(f.stride)))[:, :int(f.w) * 3]
        img = img.reshape(int(f.h), int(f.w), 3)
        box = (int(f.x0), int(f.y0), int(f.x0) + int(f.cw), int(f.y0) + int(f.ch))
        o[i] = OP.eval_frame(img, box, (rw, rh), hflip=bool(f.hflip), channel_power=pw)


def _allgather(ctx, send, recv, count, stream):
    """csrc/comm.cu dh_allgather_f32 (ncclAllGather): recv[r * count ...] = rank r's send -- over gloo here"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    src = torch.from_numpy(_f32(send, count).copy())
    dst = torch.empty(world * count, dtype=torch.float32)
    dist.all_gather_into_tensor(dst, src)
    _f32(recv, world * count)[...] = dst.numpy()


ARITHMETIC = {
    'dh_allgather_f32': _allgather, 'dh_pose_eval_f32': _pose_eval, 'dh_crop_resize_norm_u8': _crop_resize_norm,
    'dh_conv2d_f32': _conv(False), 'dh_sepconv2d_f32': _conv(True), 'dh_maxpool2d_f32': _maxpool,
    'dh_upsample2x_add_f32': _upsample_add, 'dh_add_n_f32': _add_n, 'dh_softargmax2d_f32': _softargmax2d,
    'dh_softargmax2d_ctx_f32': _softargmax2d_ctx, 'dh_softargmax3d_f32': _softargmax3d,
    'dh_softargmax3d_ex_f32': _softargmax3d_ex, 'dh_kron_pool_f32': _kron, 'dh_zeropad2d_f32': _zeropad,
    'dh_maxmin_pool2d_f32': _maxmin_pool, 'dh_global_maxmin_softmax_f32': _global_maxmin_softmax, 'dh_mask_mul_f32': _mask_mul,
}


def install(arithmetic=False):
    import torch
    import torch.distributed as dist

    from deephar_b200 import _ffi

    # ---- tensors: device memory is host memory -------------------------------------------------------------------------
    def strip_device(fn):
        def wrapper(*args, **kwargs):
            if _is_cuda(kwargs.get('device')):
                kwargs['device'] = 'cpu'
            return fn(*args, **kwargs)
        return wrapper
    for name in ('empty', 'zeros', 'ones', 'full', 'tensor', 'randn', 'rand', 'arange', 'as_tensor'):
        setattr(torch, name, strip_device(getattr(torch, name)))
    real_generator = torch.Generator
    torch.Generator = lambda device=None: real_generator()
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.record_stream = lambda self, stream: None
    torch.Tensor.is_cuda = property(lambda self: True)
    real_to = torch.Tensor.to

    def to(self, *args, **kwargs):
        args = tuple('cpu' if _is_cuda(a) and not isinstance(a, torch.dtype) else a for a in args)
        if _is_cuda(kwargs.get('device')):
            kwargs['device'] = 'cpu'
        return real_to(self, *args, **kwargs)
    torch.Tensor.to = to

    # ---- torch.cuda: one inert device ----------------------------------------------------------------------------------
    class Event(object):
        def __init__(self, enable_timing=False, **kw):
            self.t = None

        def record(self, stream=None):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return max(1e-3, (other.t - self.t) * 1000.0)

        def synchronize(self):
            pass

        def query(self):
            return True

    class Stream(object):
        cuda_stream = 0

        def __init__(self, *a, **k):
            pass

        def synchronize(self):
            pass

        def wait_event(self, event):
            pass

        def wait_stream(self, stream):
            pass

        def record_event(self, event=None):
            event = event or Event()
            event.record()
            return event

    class CUDAGraph(object):
        """records the library calls issued while it is being captured and issues them again on replay()"""

        def __init__(self):
            self.ops = []

        def replay(self):
            for fn, args in self.ops:
                fn(*args)

    @contextlib.contextmanager
    def capture(g, *a, **k):
        CAPTURING.append(g)
        try:
            yield
        finally:
            CAPTURING.pop()

    cuda = torch.cuda
    the_stream = Stream()
    cuda.is_available = lambda: True
    cuda.device_count = lambda: int(os.environ.get('FAKE_CUDA_DEVICES', '8'))
    cuda.current_device = lambda: 0
    cuda.set_device = lambda d: None
    cuda.synchronize = lambda *a, **k: None
    cuda.empty_cache = lambda: None
    cuda.current_stream = lambda *a, **k: the_stream
    cuda.Stream = Stream
    cuda.Event = Event
    cuda.CUDAGraph = CUDAGraph
    cuda.stream = lambda s: contextlib.nullcontext()
    cuda.graph = capture
    cuda.device = lambda d: contextlib.nullcontext()

    # ---- torch.distributed: NCCL -> gloo -------------------------------------------------------------------------------
    real_init = dist.init_process_group

    def init_process_group(backend=None, *args, **kwargs):
        kwargs.pop('device_id', None)
        return real_init('gloo', *args, **kwargs)
    dist.init_process_group = init_process_group
    dist.get_backend = lambda group=None: 'gloo'

    # ---- the C library: host entry points are real, device entry points are no-ops -----------------------------------
    real = _ffi.lib()

    class FakeLib(object):
        calls = {}

        def __getattr__(self, name):
            if name in HOST_ENTRY_POINTS:
                return getattr(real, name)
            if name not in _ffi.SIGNATURES:
                raise AttributeError(name)
            restype, argtypes = _ffi.SIGNATURES[name]
            proto = C.CFUNCTYPE(restype, *argtypes)             # converts / type-checks the arguments as the real call does

            def body(*args):
                FakeLib.calls[name] = FakeLib.calls.get(name, 0) + 1
                if CAPTURING:
                    CAPTURING[-1].ops.append((body, args))
                return 0
            if name == 'dh_ctx_create':
                def body(out, device):                          # noqa: F811
                    C.cast(out, C.POINTER(C.c_void_p))[0] = 0xB200
                    return 0
            if arithmetic and name in ARITHMETIC:
                op = ARITHMETIC[name]

                def body(*args, _op=op, _name=name):            # noqa: F811
                    FakeLib.calls[_name] = FakeLib.calls.get(_name, 0) + 1
                    if CAPTURING:
                        CAPTURING[-1].ops.append((body, args))
                    _op(*args)
                    return 0
            fn = proto(body)
            setattr(self, name, fn)
            return fn
    fake = FakeLib()
    _ffi._lib = fake
    _ffi.lib = lambda: fake
    return fake


def main():
    """python tests/fake_cuda.py [--arithmetic] <script.py | -m module> [args ...]"""
    argv = sys.argv[1:]
    arithmetic = bool(argv) and argv[0] == '--arithmetic'
    if arithmetic:
        argv = argv[1:]
    install(arithmetic=arithmetic)
    if argv[0] == '-m':
        sys.argv = argv[1:]
        runpy.run_module(argv[1], run_name='__main__', alter_sys=True)
    else:
        sys.argv = argv
        runpy.run_path(argv[0], run_name='__main__')


if __name__ == '__main__':
    main()
