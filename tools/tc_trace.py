#!/usr/bin/env python
"""Cycle trace of one tile of k_mlp_tc (CTA 0, thread 0): where a tile's time goes, step by step."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from vanerf_b200 import _lib as L  # noqa: E402

STEPS = ["G1", "G2", "G3", "G4", "M0", "P0", "P1", "P2", "P3", "P4", "P5", "M1", "M2", "M3", "Q1", "Q2", "Q3", "T1", "T2", "T3", "T4",
         "I1", "I2", "I3", "I4", "I5", "I6", "I7", "I8", "I9"]
n_rays = int(sys.argv[1]) if len(sys.argv) > 1 else 592          # 592 rays x 64 = 296 tiles = one full wave
H, W, V = 512, 334, 3
sc, inp, sd = parity.build_case(H, W, V, mode="ref")
r, _ = parity.make_renderer(inp, sd, "cuda:0")
tar = r.make_target(inp["cam_tar"], inp["bounds"])
ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
pix = torch.stack([xs, ys], -1).reshape(-1, 2)
sel = torch.from_numpy(np.random.RandomState(0).permutation(H * W)[:n_rays].copy())
pix = pix[sel].contiguous()
rays, z = r.sample_rays(tar, pix, 64)
geo = r.geom_query(tar, rays, z)
for _ in range(2):
    r.shade(tar, rays, z, geo, precision=L.BF16)
torch.cuda.synchronize()
buf = torch.zeros(16384, 2, dtype=torch.int64, device="cuda:0")
r.lib.dll.vanerf_tc_profile(r.ctx, buf.data_ptr(), 16384)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
r.shade(tar, rays, z, geo, precision=L.BF16)
ev1.record()
torch.cuda.synchronize()
n = r.lib.dll.vanerf_tc_profile(r.ctx, None, 0)
t = buf[:n].cpu().numpy()
print(f"shade of {n_rays * 64} samples ({n_rays * 64 // 128} tiles): {ev0.elapsed_time(ev1):.3f} ms, {n} trace points, tc_error {r.tc_error()}")
names = {1: "epi_done", 2: "fenced", 3: "issued", 4: "acc_ready", 5: "rec", 6: "tile_end", 7: "wfull", 8: "I:wait_ops", 9: "I:ops_ready"}
for who, title in ((0, "tile thread 0"), (1, "issuer of tile 0")):
    tt = [(int(tag) % 100000, int(clk)) for tag, clk in t if int(tag) // 100000 == who]
    if not tt:
        continue
    print(f"==== {title}: {len(tt)} points")
    t0 = tt[0][1]
    prev = t0
    agg = {}
    for tag, clk in tt:
        kind, st = divmod(tag, 1000)
        name = names.get(kind, "?")
        if kind == 7:
            name = "wfull_wait" if st == 0 else "wfull_ok"
        label = f"{name}:{STEPS[st] if kind in (1, 2, 3, 4, 8, 9) and st < len(STEPS) else st}"
        d = clk - prev
        agg.setdefault(name, [0, 0])
        agg[name][0] += d
        agg[name][1] += 1
        print(f"{clk - t0:9d} +{d:7d}  {label}")
        prev = clk
    print("total cycles", tt[-1][1] - t0)
    print("time attributed to the interval ENDING at each kind of marker:")
    for k, (c, m) in agg.items():
        print(f"  {k:12s} {c:9d} cycles over {m} intervals")
