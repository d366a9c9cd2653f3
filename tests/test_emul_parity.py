"""Kernel LOGIC checks without a GPU: the CUDA kernel sources compiled as plain C++ (host emulation, tests/_emul)
against the oracle.  These do not replace the `-m gpu` parity tests (same checks on the real CUDA library)."""
import numpy as np
import pytest

import parity
from vanerf_b200 import _lib as L


@pytest.mark.parametrize("H,W,V,mode,layout,npix", [
    (256, 256, 1, "ref", "narrow", 3),
    (512, 334, 3, "stress", "narrow", 4),
    (512, 334, 3, "stress", "bvv", 3),
])
def test_emulated_kernels_match_oracle(emul_lib, H, W, V, mode, layout, npix):
    sc, inp, sd = parity.build_case(H, W, V, mode=mode, layout=layout)
    r, vert_vis = parity.make_renderer(inp, sd, "cpu", emul_lib)
    pix = parity.lattice_pixels(H, W, npix)
    errs, oo, ot = parity.check_all(r, vert_vis, inp, sd, pix, precision=L.FP32)
    assert ot["valid"].any() and not ot["valid"].all(), "case must exercise both valid and invalid samples"
    assert ot["geo"]["inside"].any(), "case must contain samples inside the mesh"
    parity.check_render_rays(r, inp, oo, pix, tol_fine=5e-3)


def test_emulated_ragged_sizes(emul_lib):
    """Ray / sample counts that are not multiples of the tile sizes (64-sample MLP tile, 16-lane gather groups)."""
    import torch
    from oracle import oracle_torch as OT
    sc, inp, sd = parity.build_case(256, 256, 2, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cpu", emul_lib)
    pix = parity.lattice_pixels(256, 256, 3)[:7]
    orc = OT.Oracle(sd, inp)
    ot = {}
    orc.render(fine=False, pixels=pix, S_c=24, taps=ot)
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    rays, z = r.sample_rays(tar, torch.from_numpy(pix), 24)
    parity.assert_exact("z", z.numpy(), ot["z"])
    geo = r.geom_query(tar, rays, z)
    rgba, valid, raw = r.shade(tar, rays, z, geo)
    parity.assert_exact("valid", valid.numpy() > 0, ot["valid"])
    parity.assert_close("rgba", rgba.numpy(), ot["rgba"], parity.TOL_FP32)


def test_emulated_coarse_reuse_is_bit_identical(emul_lib):
    """vanerf_set_reuse_coarse: the fine pass evaluates only the new depths and reuses the coarse pass for the coarse
    depths inside the merged set; the rendered rows must not change by a single bit (logic check on the emulated
    kernels; the GPU suite repeats it for both precision paths)."""
    import torch
    sc, inp, sd = parity.build_case(256, 256, 2, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cpu", emul_lib)
    pix = torch.from_numpy(parity.lattice_pixels(256, 256, 3)[:7])
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    oc0, of0 = r.render_rays(tar, pix, 24, 16, True, L.FP32)
    oc0, of0 = oc0.clone(), of0.clone()
    r.set_reuse_coarse(True)
    oc1, of1 = r.render_rays(tar, pix, 24, 16, True, L.FP32)
    r.set_reuse_coarse(False)
    assert torch.equal(oc0, oc1) and torch.equal(of0, of1)
    assert torch.isfinite(of1).all() and float(of1[:, :3].abs().max()) > 0
