#!/usr/bin/env python
"""Developer aid: wall time of the per-frame setup (vanerf_frame_setup + the torch conv stacks) on the GPU box."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vanerf_b200 import synthetic, weights  # noqa: E402
from vanerf_b200.model import VANeRF  # noqa: E402

H, W, V = 512, 334, 3
dev = torch.device("cuda:0")
net = VANeRF(device=dev, precision="bf16").eval()
net.load_state_dict(weights.init_state_dict(H, W, mode="ref"))
inp = synthetic.to_torch(synthetic.make_scene(H, W, V))
mv = lambda t: t.to(dev)
img, tex, fg = mv(inp["img"]), mv(inp["feat_tex"]), mv(inp["src_foreground_mask"])
geo = [mv(t) for t in inp["feat_geo"]]
r = net.renderer
for it in range(6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gf = r.global_vertex_feature(img, tex) if hasattr(r, "global_vertex_feature") else None
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    r.set_frame(img, inp["cam_in"], inp["targets"], inp["sp_data"], geo, tex, fg)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"iter {it}: global_vertex_feature {1e3 * (t1 - t0):.2f} ms, set_frame (incl. it again) {1e3 * (t2 - t1):.2f} ms")
