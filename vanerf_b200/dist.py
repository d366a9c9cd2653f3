"""Multi-GPU sharding of one view: the reference's interleaved sub-image decomposition (src/model.py:1050-1085) used as
the ray partition across ranks, and the reassembly of the all-gathered output tiles (pixel_shuffle-style interleave).
One process per GPU; the only collective of the path is the all_gather of output tiles."""
from __future__ import annotations

import numpy as np
import torch


def grid_for_world(world: int):
    """(gy, gx) pixel periods: 1 -> 1x1, 2 -> 1x2, 4 -> 2x2, 8 -> 4x2; other sizes stripe rows."""
    return {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))


def partition_pixels(H: int, W: int, rank: int, world: int) -> np.ndarray:
    """Target pixels [x, y] (row-major over the rank's sub-grid) with (y mod gy, x mod gx) == divmod(rank, gx)."""
    gy, gx = grid_for_world(world)
    ry, rx = divmod(rank, gx)
    ys, xs = np.meshgrid(np.arange(ry, H, gy), np.arange(rx, W, gx), indexing="ij")
    return np.stack([xs, ys], -1).reshape(-1, 2).astype(np.int32)


def tile_shape(H: int, W: int, rank: int, world: int):
    gy, gx = grid_for_world(world)
    ry, rx = divmod(rank, gx)
    return len(range(ry, H, gy)), len(range(rx, W, gx))


def padded_tile_rows(H: int, W: int, world: int) -> int:
    """all_gather needs equal tile sizes: every rank pads to the largest tile."""
    return max(np.prod(tile_shape(H, W, r, world)) for r in range(world))


def assemble(tiles, H: int, W: int, world: int) -> torch.Tensor:
    """tiles[r]: (rows_r_padded, C) rows of rank r in partition_pixels order -> (H, W, C) image."""
    C = tiles[0].shape[1]
    out = torch.empty(H, W, C, dtype=tiles[0].dtype, device=tiles[0].device)
    gy, gx = grid_for_world(world)
    for r in range(world):
        ry, rx = divmod(r, gx)
        th, tw = tile_shape(H, W, r, world)
        out[ry::gy, rx::gx] = tiles[r][: th * tw].reshape(th, tw, C)
    return out
