#!/usr/bin/env python
"""Developer aid (GPU): vanerf_render_rays with geometry reuse on / off must give identical bits (both precision paths)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
from vanerf_b200 import _lib as L
sc, inp, sd = parity.build_case(512, 334, 3, mode="stress")
r, _ = parity.make_renderer(inp, sd, "cuda:0")
tar = r.make_target(inp["cam_tar"], inp["bounds"])
pix = torch.from_numpy(np.random.RandomState(1).randint(0, [334, 512], size=(3000, 2)).astype(np.int32))
for prec in (L.FP32, L.BF16):
    outs = []
    for on in (True, False):
        r.set_reuse_geometry(on)
        oc, of = r.render_rays(tar, pix, 64, 64, True, prec)
        outs.append((oc.cpu().numpy(), of.cpu().numpy()))
    print("precision", prec, "coarse equal", np.array_equal(outs[0][0], outs[1][0]), "fine equal", np.array_equal(outs[0][1], outs[1][1]),
          "tc_error", r.tc_error(), "fine finite", np.isfinite(outs[0][1]).all())
