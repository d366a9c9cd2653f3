import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


@pytest.fixture(scope="session")
def emul_lib():
    """Host-emulation build of the kernel sources (logic checks without a GPU).  Never used by the product."""
    from tools import build as B
    from vanerf_b200 import _lib as L
    return L.Lib(B.build_emul(), emulated=True)


@pytest.fixture(scope="session")
def cuda_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from vanerf_b200 import _lib as L
    return L.get_lib()
