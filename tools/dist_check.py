#!/usr/bin/env python
"""torchrun helper of tests/test_gpu_parity.py::test_cuda_two_rank_nccl_image_equals_single_gpu_image: every rank renders its
interleaved tile of one view, the tiles are all-gathered over NCCL (vanerf_b200.dist.render_view) and rank 0 compares the
assembled image with the single-GPU render, bit for bit, on both precision paths."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from vanerf_b200 import _lib as L  # noqa: E402
from vanerf_b200 import dist as D  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W, V = 128, 96, 3
sc, inp, sd = parity.build_case(H, W, V, mode="stress")
r, _ = parity.make_renderer(inp, sd, f"cuda:{local}")
tar = r.make_target(inp["cam_tar"], inp["bounds"])
ok = True
for prec in (L.FP32, L.BF16):
    img = D.render_view(r, tar, H, W, rank, world, 32, 32, True, prec)
    r.finish()
    if rank == 0:
        full = D.render_view(r, tar, H, W, 0, 1, 32, 32, True, prec)
        same = torch.equal(img, full)
        print(f"precision {prec}: {world}-rank image == 1-GPU image: {same}", flush=True)
        ok = ok and same and bool(torch.isfinite(img).all())
dist.barrier()
if rank == 0:
    print("DIST_CHECK_OK" if ok else "DIST_CHECK_FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
