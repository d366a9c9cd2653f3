#!/usr/bin/env python
"""Short single-GPU run of the render path for ncu: one 334x512 frame setup + `--rays` rays (coarse + fine)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=8192)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--iters", type=int, default=2)
    args = ap.parse_args()
    import parity
    from vanerf_b200 import _lib as L
    H, W, V = 512, 334, 3
    sc, inp, sd = parity.build_case(H, W, V, mode="ref")
    r, _ = parity.make_renderer(inp, sd, "cuda:0")
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pix = torch.stack([xs, ys], -1).reshape(-1, 2)
    sel = torch.from_numpy(np.random.RandomState(0).permutation(H * W)[:args.rays].copy())
    pix = pix[sel].contiguous().to("cuda:0")
    prec = L.FP32 if args.precision == "fp32" else L.BF16
    for _ in range(args.iters):
        oc, of = r.render_rays(tar, pix, 64, 64, True, prec)
    torch.cuda.synchronize()
    print("ok", float(of[:, :3].abs().mean()), "launches", r.launches)


if __name__ == "__main__":
    main()
