#!/usr/bin/env python
"""Developer aid: torch.profiler table of one training step (workload E): device time by kernel, CPU vs device totals."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vanerf_b200 import synthetic, weights  # noqa: E402
from vanerf_b200 import train as T  # noqa: E402

H, W, V, patch = 512, 334, 3, 64
dev = torch.device("cuda", 0)
inp = synthetic.to_torch(synthetic.make_scene(H, W, V), dev)
path = T.TrainableRenderPath(weights.init_state_dict(H, W, mode="ref"), dev, rand_noise_std=0.01)
opt = torch.optim.Adam(path.parameters(), lr=1e-3)
path.set_frame(inp)
msk = inp["src_foreground_mask"][0, 0, 0].bool().cpu()
rand = T.TrainRandom(1000, 2000)
target = torch.rand(patch * patch, 3).to(dev)
cfg = dict(training=True, uniform=False, fine=True, S_c=64, S_f=64)


def step():
    pix = T.patch_pixels(msk, W, H, patch, patch, rand)
    return T.training_step(path, inp, pix, target, opt, rand, 1, **cfg)


for _ in range(2):
    step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
step()
torch.cuda.synchronize()
print(f"wall per step {1e3 * (time.perf_counter() - t0):.1f} ms")
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
