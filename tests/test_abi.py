"""The C-ABI library loads and exports every symbol include/vanerf_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "vanerf_b200.h")).read()
    return sorted(set(re.findall(r"\b(vanerf_[a-z0-9_]+)\s*\(", h)))


def test_cuda_library_exports_every_declared_symbol():
    from tools import build as B
    from vanerf_b200 import _lib as L
    so = B.build_cuda()
    dll = ctypes.CDLL(so)
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/vanerf_b200.h but not exported"
    assert set(L.PROTOTYPES) == set(names), set(L.PROTOTYPES) ^ set(names)


def test_product_path_fails_loudly_without_the_extension(tmp_path):
    from vanerf_b200 import _lib as L
    with pytest.raises(L.VanerfError):
        L.Lib(str(tmp_path / "libvanerf_b200.so"))


def test_product_refuses_cpu_device():
    import torch
    from vanerf_b200 import _lib as L
    from vanerf_b200.renderer import Renderer
    if torch.cuda.is_available():
        pytest.skip("GPU box")
    from tools import build as B
    B.build_cuda()
    with pytest.raises(L.VanerfError):
        Renderer("cpu")                  # CUDA library + cpu device: no CPU fallback


def test_status_strings_and_argument_errors_without_gpu():
    from tools import build as B
    dll = ctypes.CDLL(B.build_cuda())
    dll.vanerf_status_str.restype = ctypes.c_char_p
    assert dll.vanerf_status_str(0) == b"ok"
    assert b"invalid" in dll.vanerf_status_str(-1)
    assert dll.vanerf_ctx_create(None, 0) == -1
