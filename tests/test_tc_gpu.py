"""Tensor-core (tcgen05 / TMEM / bulk-copy) path on a B200: the MMA building block against torch, then the fused bf16
shading kernel against the oracle within the bf16 bar of BASELINE.json (1e-2 max-abs)."""
import numpy as np
import pytest
import torch

import parity
from vanerf_b200 import _lib as L
from vanerf_b200.renderer import Renderer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K,N", [(16, 16), (64, 64), (48, 96), (128, 128), (208, 16), (256, 120), (144, 33)])
def test_tc_mma_block_matches_torch(cuda_lib, K, N):
    r = Renderer("cuda:0")
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = torch.randn(128, K, generator=g)
    W = torch.randn(N, K, generator=g)
    D = r.tc_selftest(A, W).cpu()
    ref = A.bfloat16().float() @ W.bfloat16().float().T
    err = (D - ref).abs().max().item()
    assert err < 2e-4 * max(1.0, ref.abs().max().item()), f"K={K} N={N}: max-abs {err:.3e}"


@pytest.mark.parametrize("H,W,V,mode,layout,npix", [
    (256, 256, 1, "ref", "narrow", 8),
    (256, 256, 1, "stress", "narrow", 8),
    (512, 334, 3, "ref", "narrow", 10),
    (512, 334, 3, "stress", "narrow", 12),
    (512, 334, 3, "stress", "bvv", 12),
    (512, 334, 2, "stress", "bvv", 6),
])
def test_tc_shading_matches_oracle_bf16(cuda_lib, H, W, V, mode, layout, npix):
    sc, inp, sd = parity.build_case(H, W, V, mode=mode, layout=layout)
    r, vert_vis = parity.make_renderer(inp, sd, "cuda:0")
    pix = parity.lattice_pixels(H, W, npix)
    rep = {}
    try:
        errs, oo, ot = parity.check_all(r, vert_vis, inp, sd, pix, precision=L.BF16, report=rep)
    finally:
        print("bf16 max-abs errors so far", {k: f"{v:.2e}" for k, v in rep.items()}, "tc_error", r.tc_error())
    assert r.tc_error() == 0
    parity.check_render_rays(r, inp, oo, pix, precision=L.BF16)
    print("bf16 max-abs errors", {k: f"{v:.2e}" for k, v in errs.items()})


def test_tc_rejects_more_than_three_views(cuda_lib):
    sc, inp, sd = parity.build_case(256, 256, 4, mode="ref")
    r, _ = parity.make_renderer(inp, sd, "cuda:0")
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    with pytest.raises(L.VanerfError):
        r.render_rays(tar, torch.from_numpy(parity.lattice_pixels(256, 256, 4)), 64, 64, True, L.BF16)


def test_tc_full_view_bf16_properties(cuda_lib):
    """The HEADLINE configuration (BASELINE.json configs[1]: 334x512 view, 171 008 rays, V=3, 64 + 128 evaluations per ray) on
    the bf16 tensor-core path: finite everywhere, chunk / launch independence (rows of the chunked full-view call equal the same
    rays rendered as one small batch, bit for bit) and an oracle spot check on 96 random rays at the literal north-star bar."""
    from oracle import oracle_torch as OT
    H, W, V = 512, 334, 3
    sc, inp, sd = parity.build_case(H, W, V, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cuda:0")
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pix = torch.stack([xs, ys], -1).reshape(-1, 2)
    R = pix.shape[0]
    assert R == 171008
    oc, of = r.render_rays(tar, pix, 64, 64, True, L.BF16)
    r.finish()                                   # raises if a tensor-core launch gave up on a bounded wait
    oc, of = oc.cpu().numpy(), of.cpu().numpy()
    assert np.isfinite(oc).all() and np.isfinite(of).all()
    for a in (oc[:, 4], of[:, 4]):
        assert a.max() < 1 + 1e-3 and a.min() > 0.95
    sel = np.random.RandomState(0).choice(R, 96, replace=False)
    oc2, of2 = r.render_rays(tar, pix[sel], 64, 64, True, L.BF16)
    r.finish()
    assert np.array_equal(oc2.cpu().numpy(), oc[sel]) and np.array_equal(of2.cpu().numpy(), of[sel]), "result depends on the ray batch"
    oo = OT.Oracle(sd, inp).render(fine=True, pixels=pix[sel].numpy())
    e = {"tex_fg": parity.assert_close("full-view bf16 tex_fg vs oracle", oc[sel, :3], oo["tex_fg"], parity.TOL_BF16),
         "alpha": parity.assert_close("full-view bf16 alpha vs oracle", oc[sel, 4], oo["alpha"], parity.TOL_BF16),
         "tex_fg_fine": parity.assert_close("full-view bf16 tex_fg_fine vs oracle (own fine depths)", of[sel, :3], oo["tex_fg_fine"],
                                            parity.TOL_E2E_FINE_BF16),
         "alpha_fine": parity.assert_close("full-view bf16 alpha_fine vs oracle", of[sel, 4], oo["alpha_fine"], parity.TOL_BF16)}
    print("full-view bf16 max-abs errors", {k: f"{v:.2e}" for k, v in e.items()})
