"""Bounds-checked developer build of the tensor-core gather (-DGTC_DEBUG: every map / table / sample index the kernel forms is
checked against its array and reported by device printf).  compute-sanitizer is not available on the GPU pool, so this is the
memory-safety evidence for the gather kernel: the build must stay compilable and must report nothing on the full-size frame."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import torch
import parity
from vanerf_b200 import _lib as L
sc, inp, sd = parity.build_case(512, 334, 3, mode="stress", layout="bvv")
r, _ = parity.make_renderer(inp, sd, "cuda:0")
tar = r.make_target(inp["cam_tar"], inp["bounds"])
pix = torch.from_numpy(parity.lattice_pixels(512, 334, 24))              # 576 rays incl. image borders
oc, of = r.render_rays(tar, pix, 64, 64, True, L.BF16)
torch.cuda.synchronize()
r.finish()
print("DEBUG_BUILD_RAN", bool(torch.isfinite(of).all()))
'''


@pytest.mark.gpu
def test_bounds_checked_gather_build_reports_nothing(cuda_lib):
    out_dir = os.path.join(HERE, "_dbg")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libvanerf_b200_dbg.so")
    src = os.path.join(ROOT, "vanerf_b200", "csrc", "vanerf_b200.cu")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-DGTC_DEBUG",
                           "-Xcompiler", "-fPIC", "-shared", "-o", so, src])
    env = dict(os.environ, VANERF_B200_LIB=so)
    p = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT, "tests": HERE}], env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "DEBUG_BUILD_RAN True" in p.stdout
    assert "GTC_DEBUG" not in p.stdout, p.stdout[:2000]
