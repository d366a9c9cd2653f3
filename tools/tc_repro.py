#!/usr/bin/env python
"""Small reproducer runs of the tensor-core path (for compute-sanitizer / debugging)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from vanerf_b200 import _lib as L  # noqa: E402

H, W, V, mode, npix, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
sc, inp, sd = parity.build_case(H, W, V, mode=mode)
r, _ = parity.make_renderer(inp, sd, "cuda:0")
tar = r.make_target(inp["cam_tar"], inp["bounds"])
pix = torch.from_numpy(parity.lattice_pixels(H, W, npix))
for i in range(reps):
    oc, of = r.render_rays(tar, pix, 64, 64, True, L.BF16)
    torch.cuda.synchronize()
    print(i, "tc_error", r.tc_error(), "finite", bool(torch.isfinite(of).all()), float(of[:, :3].abs().mean()), flush=True)
