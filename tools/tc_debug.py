#!/usr/bin/env python
"""Diagnostics of the tensor-core path on a GPU box: MMA building block vs torch, then per-stage errors of the bf16
shading kernel vs the oracle (no asserts: prints everything it can)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from oracle import oracle_torch as OT  # noqa: E402
from vanerf_b200 import _lib as L  # noqa: E402
from vanerf_b200.renderer import Renderer  # noqa: E402


def selftests():
    r = Renderer("cuda:0")
    for K, N in [(16, 16), (64, 64), (64, 128), (128, 128), (256, 128), (208, 16), (48, 96), (144, 33)]:
        g = torch.Generator().manual_seed(K * 1000 + N)
        A = torch.randn(128, K, generator=g)
        W = torch.randn(N, K, generator=g)
        try:
            D = r.tc_selftest(A, W).cpu()
        except Exception as e:  # noqa: BLE001
            print(f"selftest K={K} N={N}: FAILED {e}")
            continue
        ref = A.bfloat16().float() @ W.bfloat16().float().T
        err = (D - ref).abs()
        print(f"selftest K={K:3d} N={N:3d}: max-abs err {err.max().item():.3e} (|ref| max {ref.abs().max().item():.2f})", flush=True)
        if err.max().item() > 1e-2:
            bad_rows = (err.max(1).values > 1e-2).nonzero().flatten().tolist()
            bad_cols = (err.max(0).values > 1e-2).nonzero().flatten().tolist()
            print("   bad rows", bad_rows[:16], "... n=", len(bad_rows), " bad cols", bad_cols[:16], "... n=", len(bad_cols))
            print("   D[0,:8]", D[0, :8].tolist(), "\n   ref[0,:8]", ref[0, :8].tolist())


def shading(H, W, V, mode, npix, layout="narrow"):
    sc, inp, sd = parity.build_case(H, W, V, mode=mode, layout=layout)
    r, vert_vis = parity.make_renderer(inp, sd, "cuda:0")
    pix = parity.lattice_pixels(H, W, npix)
    orc = OT.Oracle(sd, inp)
    ot = {}
    orc.render(fine=False, pixels=pix, S_c=64, taps=ot)
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    rays, z = r.sample_rays(tar, torch.from_numpy(pix), 64)
    geo = r.geom_query(tar, rays, z)
    t0 = time.time()
    rgba, valid, raw, lat = r.shade(tar, rays, z, geo, precision=L.BF16, want_latent=True)
    code = r.tc_error()
    print(f"shade bf16 {H}x{W} V={V} {mode}: tc_error={code} ({time.time() - t0:.2f}s)", flush=True)
    rgba32, valid32, raw32, lat32 = r.shade(tar, rays, z, geo, precision=L.FP32, want_latent=True)
    v = ot["valid"].astype(bool)
    ref_raw = np.concatenate([ot["query"]["o"], ot["query"]["rgb"]], 1)
    for name, got, ref in [("latent", lat, ot["query"]["latent"]), ("raw", raw, ref_raw), ("rgba", rgba, ot["rgba"])]:
        g = got.cpu().numpy()
        e = np.abs(g - ref)
        fin = np.isfinite(g).all()
        print(f"   {name:7s} finite={fin} max-abs err all={np.nanmax(e):.3e} valid-only={np.nanmax(e[v]) if v.any() else 0:.3e} "
              f"per-col {np.array2string(np.nanmax(e, 0)[:8], precision=2)} |ref| max {np.abs(ref).max():.3f}")
    print("   valid equal:", bool((valid.cpu().numpy() > 0).__eq__(v).all()), " n_valid", int(v.sum()), "/", v.size)
    e32 = np.abs(raw32.cpu().numpy() - ref_raw).max()
    print(f"   (fp32 path raw err {e32:.3e})")


if __name__ == "__main__":
    selftests()
    torch.cuda.synchronize()
    for cfg in [(256, 256, 1, "ref", 6), (256, 256, 1, "stress", 6), (512, 334, 3, "ref", 8), (512, 334, 3, "stress", 8)]:
        try:
            shading(*cfg)
        except Exception as e:  # noqa: BLE001
            print("shading", cfg, "FAILED:", repr(e)[:400], flush=True)
