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


def _is_cuda(dev):
    return dev is not None and str(dev).startswith('cuda')


def install():
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
        def replay(self):
            pass

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
    cuda.graph = lambda g, *a, **k: contextlib.nullcontext()
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
                return 0
            if name == 'dh_ctx_create':
                def body(out, device):                          # noqa: F811
                    C.cast(out, C.POINTER(C.c_void_p))[0] = 0xB200
                    return 0
            fn = proto(body)
            setattr(self, name, fn)
            return fn
    fake = FakeLib()
    _ffi._lib = fake
    _ffi.lib = lambda: fake
    return fake


def main():
    install()
    script = sys.argv[1]
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
