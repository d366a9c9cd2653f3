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


def render_view(renderer, tar, H: int, W: int, rank: int, world: int, n_coarse=64, n_fine=64, fine=True, precision=0, group=None,
                gather=None) -> torch.Tensor:
    """One view on `world` ranks: this rank renders its interleaved pixel subset (`partition_pixels`), the (rows, 16) output
    tiles [fine r,g,b,depth,alpha,sdf,0,0 | coarse r,g,b,depth,alpha,sdf,0,0] are all-gathered (the path's only collective) and
    interleaved back into the (H, W, 16) image in the reference's pixel order.  Every rank returns the full image; it is
    bit-identical to the single-GPU render because a ray's result does not depend on its batch.
    `gather(tile) -> list of world tiles` replaces torch.distributed.all_gather (tests run the ranks one after the other)."""
    pix = torch.from_numpy(partition_pixels(H, W, rank, world))
    oc, of = renderer.render_rays(tar, pix, n_coarse, n_fine, fine, precision)
    rows = padded_tile_rows(H, W, world)
    tile = oc.new_zeros((rows, 16))
    tile[: oc.shape[0], 8:] = oc
    if fine:
        tile[: of.shape[0], :8] = of
    if world == 1 and gather is None:
        tiles = [tile]
    elif gather is not None:
        tiles = gather(tile)
    else:
        import torch.distributed as dist
        tiles = [torch.empty_like(tile) for _ in range(world)]
        dist.all_gather(tiles, tile, group=group)
    return assemble(tiles, H, W, world)
