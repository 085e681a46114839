"""The C-ABI library loads and exports every symbol include/deephar_b200.h declares
(no compute calls: there is no GPU in the CPU test run)."""
import ctypes
import os
import re

from deephar_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'deephar_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dh_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_all_declared_symbols():
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), 'missing export %s' % n


def test_binding_covers_header():
    assert sorted(_ffi.SIGNATURES) == _declared()


def test_version_and_error_string():
    lib = _ffi.lib()
    assert lib.dh_version() >= 100
    assert isinstance(lib.dh_last_error(), bytes)


def test_no_gpu_fails_loudly():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from deephar_b200 import reception
    m = reception.build((64, 64, 3), 16, 2, num_blocks=1).init_synthetic_weights()
    import numpy as np
    with pytest.raises(_ffi.DeepharB200Error):
        m.predict(np.zeros((1, 64, 64, 3), np.float32))
    h = ctypes.c_void_p()
    assert lib_rc(h) != 0


def lib_rc(h):
    return _ffi.lib().dh_ctx_create(ctypes.byref(h), 0)
