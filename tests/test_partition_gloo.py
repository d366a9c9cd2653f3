"""N > 1 host logic on CPU: world-size-2 (and 4) gloo groups partition a view, all_gather tiles and reassemble it."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vanerf_b200 import dist as vd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, H, W, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pix = vd.partition_pixels(H, W, rank, world)
    rows = vd.padded_tile_rows(H, W, world)
    # stand-in for the render: a function of the pixel only (what matters here is routing, not shading)
    tile = torch.zeros(rows, 3)
    tile[: pix.shape[0], 0] = torch.from_numpy(pix[:, 0].astype(np.float32))
    tile[: pix.shape[0], 1] = torch.from_numpy(pix[:, 1].astype(np.float32))
    tile[: pix.shape[0], 2] = rank
    tiles = [torch.empty_like(tile) for _ in range(world)]
    dist.all_gather(tiles, tile)
    img = vd.assemble(tiles, H, W, world)
    if rank == 0:
        q.put(img.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,H,W", [(2, 512, 334), (4, 31, 17), (2, 7, 5)])
def test_interleaved_partition_roundtrip(world, H, W):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, H, W, q)) for r in range(world)]
    for p in procs:
        p.start()
    img = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    assert np.array_equal(img[..., 0], xs) and np.array_equal(img[..., 1], ys)
    gy, gx = vd.grid_for_world(world)
    assert np.array_equal(img[..., 2], (ys % gy) * gx + (xs % gx))


def test_partition_is_a_disjoint_cover():
    for world in (1, 2, 4, 8, 3):
        H, W = 512, 334
        seen = np.zeros((H, W), int)
        for r in range(world):
            p = vd.partition_pixels(H, W, r, world)
            seen[p[:, 1], p[:, 0]] += 1
        assert (seen == 1).all()
