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


# ---------------------------------------------------------------- render_dynamic: frames dealt round-robin to the ranks
def _frame_worker(rank, world, port, n_frames, q):
    from vanerf_b200 import dynamic as vdyn
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = vdyn.frames_for_rank(n_frames, rank, world)
    # stand-in for the render of frame f: an image that encodes f and the rank that produced it
    local = [torch.stack([torch.full((2, 3), float(f)), torch.full((2, 3), float(rank))]) for f in ids]
    frames = vdyn.gather_frames(local, n_frames, rank, world)
    if rank == 0:
        q.put(torch.stack(frames).numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_frames", [(2, 5), (2, 4), (4, 6)])
def test_frame_round_robin_gather(world, n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_frame_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert out.shape == (n_frames, 2, 2, 3)
    for f in range(n_frames):
        assert (out[f, 0] == f).all() and (out[f, 1] == f % world).all()


def test_frame_partition_is_a_disjoint_cover():
    from vanerf_b200 import dynamic as vdyn
    for world in (1, 2, 4, 8):
        for n in (8, 64, 13):
            got = sorted(f for r in range(world) for f in vdyn.frames_for_rank(n, r, world))
            assert got == list(range(n))
            sizes = [len(vdyn.frames_for_rank(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def test_orbit_cameras_layout():
    from vanerf_b200 import dynamic as vdyn
    K = torch.tensor([[1100.0, 0, 167.0], [0, 1100.0, 256.0], [0, 0, 1]])
    cams = vdyn.orbit_cameras(8, K)
    assert len(cams) == 8
    for c in cams:
        assert c["K"].shape == (1, 4, 4) and c["RT"].shape == (1, 4, 4) and c["KRT"].shape == (1, 4, 4)
        R = c["RT"][0, :3, :3]
        assert torch.allclose(R @ R.T, torch.eye(3), atol=1e-5)
        # camera centre on the unit sphere, looking at the origin: the origin projects to the principal point
        o = c["KRT"][0, :3, 3]
        assert abs(o[0] / o[2] - 167.0) < 1e-3 and abs(o[1] / o[2] - 256.0) < 1e-3
