#!/usr/bin/env python
"""Developer aid: distribution of the per-sample errors of the bf16 shading kernel against the oracle (coarse and fine pass taps):
rms, percentiles, max, share of samples above 1e-2 / 2e-2 of the range.  usage: tc_errstats.py [layout] [npix]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from oracle import oracle_torch as OT  # noqa: E402
from vanerf_b200 import _lib as L  # noqa: E402

layout = sys.argv[1] if len(sys.argv) > 1 else "bvv"
npix = int(sys.argv[2]) if len(sys.argv) > 2 else 12
H, W, V = 512, 334, 3
sc, inp, sd = parity.build_case(H, W, V, mode="stress", layout=layout)
r, vert_vis = parity.make_renderer(inp, sd, "cuda:0")
pix = parity.lattice_pixels(H, W, npix)
orc = OT.Oracle(sd, inp)
ot = {}
oo = orc.render(fine=True, pixels=pix, S_c=64, S_f=64, taps=ot)
tar = r.make_target(inp["cam_tar"], inp["bounds"])
rays, z = r.sample_rays(tar, torch.from_numpy(pix), 64)
for name, zz, ref in (("coarse", z, ot["rgba"]), ("fine", torch.from_numpy(ot["z_fine"]).to(r.device), ot["rgba_fine"])):
    geo = r.geom_query(tar, rays, zz)
    rgba, valid, raw = r.shade(tar, rays, zz, geo, precision=L.BF16)
    e = np.abs(rgba.cpu().numpy()[:, 2:] - ref[:, 2:]).ravel()
    rng = float(np.abs(ref).max())
    print(f"{layout} {name}: n={e.size} range {rng:.3f} rms {np.sqrt((e ** 2).mean()):.3e} p99 {np.percentile(e, 99):.3e} p99.9 {np.percentile(e, 99.9):.3e} "
          f"p99.99 {np.percentile(e, 99.99):.3e} max {e.max():.3e} = {e.max() / rng:.3e} of range; >1e-2*range: {(e > 1e-2 * rng).sum()}, >2e-2*range: {(e > 2e-2 * rng).sum()}")
    comp = r.composite(rgba, zz, geo["sdf"].view(zz.shape))
    key = "tex_fg" if name == "coarse" else "tex_fg_fine"
    ep = np.abs(comp["color"].cpu().numpy() - oo[key])
    print(f"   per pixel {key}: max {ep.max():.3e} rms {np.sqrt((ep ** 2).mean()):.3e}")
