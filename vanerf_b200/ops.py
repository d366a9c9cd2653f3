"""torch custom-op layer over the C ABI ("thin C-ABI torch custom-op layer").

Every stage entry point of libvanerf_b200.so is registered as a `torch.library.custom_op` in the `vanerf_b200`
namespace — `sample_rays`, `geom_query`, `shade`, `composite`, `importance_sample`, `query_points`, `render_rays` —
with fake (meta) implementations, so that the path is visible to `torch.ops`, traceable, and carries shapes / dtypes
without touching the device.  `vanerf_b200.renderer.Renderer` (and through it `vanerf_b200.model.VANeRF`) calls the
kernels only through these ops.

Conventions: `ctx` is the integer handle of a live `Renderer` (`Renderer.handle`; a `vanerf_ctx` cannot cross the
op boundary as a pointer type); `tar` is the 29-float CPU tensor form of `vanerf_target` (`pack_target`): inv_K (9),
R (9), cam_pos (3), znear, zfar, bounds (6).  All other tensors are device tensors owned by the caller; the ops
enqueue on the current stream of the renderer's device and do not synchronise.  There is no CPU implementation: an op
called with a handle whose library is missing raises.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Tuple

import torch
from torch.library import custom_op

from . import _lib as L

_REGISTRY: "weakref.WeakValueDictionary[int, object]" = weakref.WeakValueDictionary()
_next_handle = [1]


def register_renderer(r) -> int:
    h = _next_handle[0]
    _next_handle[0] += 1
    _REGISTRY[h] = r
    return h


def _renderer(h: int):
    r = _REGISTRY.get(h)
    if r is None:
        raise L.VanerfError(f"vanerf_b200 op called with a dead renderer handle ({h})")
    return r


def pack_target(t: L.VTarget) -> torch.Tensor:
    return torch.tensor(list(t.inv_K) + list(t.R) + list(t.cam_pos) + [t.znear, t.zfar] + list(t.bounds), dtype=torch.float32)


def unpack_target(x: torch.Tensor) -> L.VTarget:
    v = x.detach().cpu().float().reshape(-1).tolist()
    assert len(v) == 29, "tar: 29 floats (inv_K 9, R 9, cam_pos 3, znear, zfar, bounds 6)"
    t = L.VTarget()
    t.inv_K[:] = v[0:9]
    t.R[:] = v[9:18]
    t.cam_pos[:] = v[18:21]
    t.znear, t.zfar = v[21], v[22]
    t.bounds[:] = v[23:29]
    return t


def _p(r, t):
    return r._ptr(t)


def _chk(r, st, what):
    r.lib.check(r.ctx, st, what)


# ------------------------------------------------------------------------------------------------ sample_rays
@custom_op("vanerf_b200::sample_rays", mutates_args=())
def sample_rays(ctx: int, tar: torch.Tensor, pix_xy: torch.Tensor, ztab: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Ray generation + box clip + depths (src/model.py:1190-1238, :1497-1570).  pix_xy (R,2) int32, ztab (S) or (R,S) = the
    interpolation parameters t in z = near + (far - near) t -> rays (R,8), z (R,S)."""
    r = _renderer(ctx)
    R, S = pix_xy.shape[0], ztab.shape[-1]
    rays, z = r.empty((R, L.RAY_STRIDE)), r.empty((R, S))
    t = unpack_target(tar)
    if ztab.dim() == 1:
        st = r.lib.dll.vanerf_sample_rays(r.ctx, C.byref(t), _p(r, pix_xy), R, _p(r, ztab), S, _p(r, rays), _p(r, z), r.stream)
    else:
        st = r.lib.dll.vanerf_sample_rays_t(r.ctx, C.byref(t), _p(r, pix_xy), R, _p(r, ztab.contiguous()), S, 1, _p(r, rays), _p(r, z), r.stream)
    _chk(r, st, "vanerf_sample_rays")
    return rays, z


@sample_rays.register_fake
def _(ctx, tar, pix_xy, ztab):
    R, S = pix_xy.shape[0], ztab.shape[-1]
    return pix_xy.new_empty((R, L.RAY_STRIDE), dtype=torch.float32), pix_xy.new_empty((R, S), dtype=torch.float32)


# ------------------------------------------------------------------------------------------------ geom_query
@custom_op("vanerf_b200::geom_query", mutates_args=())
def geom_query(ctx: int, tar: torch.Tensor, rays: torch.Tensor, z: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """cal_vis_sdf_batch + knn_points(K=1) (mesh_util.py:498-524, networks.py:28) -> pts (N,3), sdf (N), closest face (N) int32,
    nearest vertex (N) int32, per-view sample visibility (V,N) uint8; N = R*S, sample index fastest."""
    r = _renderer(ctx)
    R, S = z.shape
    N, V = R * S, r.frame["V"]
    pts, sdf = r.empty((N, 3)), r.empty((N,))
    face, nn = r.empty((N,), torch.int32), r.empty((N,), torch.int32)
    qvis = r.empty((V, N), torch.uint8)
    t = unpack_target(tar)
    st = r.lib.dll.vanerf_geom_query(r.ctx, C.byref(t), _p(r, rays), _p(r, z), R, S, _p(r, pts), _p(r, sdf), _p(r, face), _p(r, nn), _p(r, qvis), r.stream)
    _chk(r, st, "vanerf_geom_query")
    return pts, sdf, face, nn, qvis


@geom_query.register_fake
def _(ctx, tar, rays, z):
    R, S = z.shape
    N = R * S
    V = _renderer(ctx).frame["V"]
    return (z.new_empty((N, 3)), z.new_empty((N,)), z.new_empty((N,), dtype=torch.int32), z.new_empty((N,), dtype=torch.int32),
            z.new_empty((V, N), dtype=torch.uint8))


# ------------------------------------------------------------------------------------------------ shade
@custom_op("vanerf_b200::shade", mutates_args=())
def shade(ctx: int, tar: torch.Tensor, rays: torch.Tensor, z: torch.Tensor, sdf: torch.Tensor, nn: torch.Tensor, qvis: torch.Tensor,
          precision: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """VANeRF.query + eval_func (src/model.py:748-957, :1140-1160) -> rgba (N,5), valid (N) uint8, raw query output (N,5)."""
    r = _renderer(ctx)
    R, S = z.shape
    N = R * S
    rgba, valid, raw = r.empty((N, 5)), r.empty((N,), torch.uint8), r.empty((N, 5))
    t = unpack_target(tar)
    st = r.lib.dll.vanerf_shade(r.ctx, precision, C.byref(t), _p(r, rays), _p(r, z), R, S, _p(r, sdf), _p(r, nn), _p(r, qvis),
                                _p(r, rgba), _p(r, valid), _p(r, raw), r.stream)
    _chk(r, st, "vanerf_shade")
    return rgba, valid, raw


@shade.register_fake
def _(ctx, tar, rays, z, sdf, nn, qvis, precision):
    N = z.shape[0] * z.shape[1]
    return z.new_empty((N, 5)), z.new_empty((N,), dtype=torch.uint8), z.new_empty((N, 5))


# ------------------------------------------------------------------------------------------------ composite
@custom_op("vanerf_b200::composite", mutates_args=())
def composite(ctx: int, rgba: torch.Tensor, z: torch.Tensor, mesh_sdf: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """VANeRF.rgba2out (src/model.py:1465-1494) -> color (R,3), depth (R), alpha (R), sdf (R), contrib (R,S)."""
    r = _renderer(ctx)
    R, S = z.shape
    color, depth, alpha, sdf, contrib = r.empty((R, 3)), r.empty((R,)), r.empty((R,)), r.empty((R,)), r.empty((R, S))
    st = r.lib.dll.vanerf_composite(r.ctx, _p(r, rgba), _p(r, z), _p(r, mesh_sdf), R, S, _p(r, color), _p(r, depth), _p(r, alpha),
                                    _p(r, sdf), _p(r, contrib), r.stream)
    _chk(r, st, "vanerf_composite")
    return color, depth, alpha, sdf, contrib


@composite.register_fake
def _(ctx, rgba, z, mesh_sdf):
    R, S = z.shape
    return z.new_empty((R, 3)), z.new_empty((R,)), z.new_empty((R,)), z.new_empty((R,)), z.new_empty((R, S))


# ------------------------------------------------------------------------------------------------ importance_sample
@custom_op("vanerf_b200::importance_sample", mutates_args=())
def importance_sample(ctx: int, contrib: torch.Tensor, z: torch.Tensor, u: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """VANeRF.importance_sample + sort(cat[z, z_fine]) (src/model.py:1301-1307, :1425-1462).  u (n_fine) shared by all rays
    (linspace for uniform=True) or (R,n_fine) per ray (training: torch.rand) -> z_fine (R,n_fine), z_all (R,S+n_fine) sorted."""
    r = _renderer(ctx)
    R, S = z.shape
    nf = u.shape[-1]
    z_f, z_all = r.empty((R, nf)), r.empty((R, S + nf))
    st = r.lib.dll.vanerf_importance(r.ctx, _p(r, contrib), _p(r, z), R, S, _p(r, u), nf, 1 if u.dim() == 2 else 0, _p(r, z_f), _p(r, z_all), r.stream)
    _chk(r, st, "vanerf_importance")
    return z_f, z_all


@importance_sample.register_fake
def _(ctx, contrib, z, u):
    R, S = z.shape
    return z.new_empty((R, u.shape[-1])), z.new_empty((R, S + u.shape[-1]))


# ------------------------------------------------------------------------------------------------ query_points
@custom_op("vanerf_b200::query_points", mutates_args=())
def query_points(ctx: int, tar: torch.Tensor, pts: torch.Tensor, view: torch.Tensor, query_sdf: torch.Tensor, query_vis: torch.Tensor,
                 precision: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """VANeRF.query on explicit points (src/model.py:748-877).  query_sdf (N) / query_vis (V,N) uint8 with numel() == 0 mean
    "compute from the mesh" -> raw (N,5) = [o0,o1,r,g,b], valid (N) uint8, rgba (N,5)."""
    r = _renderer(ctx)
    N = pts.shape[0]
    raw, valid, rgba = r.empty((N, 5)), r.empty((N,), torch.uint8), r.empty((N, 5))
    t = unpack_target(tar)
    st = r.lib.dll.vanerf_query_points(r.ctx, precision, C.byref(t), _p(r, pts), _p(r, view), N,
                                       _p(r, query_sdf) if query_sdf.numel() else None, _p(r, query_vis) if query_vis.numel() else None,
                                       _p(r, raw), _p(r, valid), _p(r, rgba), r.stream)
    _chk(r, st, "vanerf_query_points")
    return raw, valid, rgba


@query_points.register_fake
def _(ctx, tar, pts, view, query_sdf, query_vis, precision):
    N = pts.shape[0]
    return pts.new_empty((N, 5)), pts.new_empty((N,), dtype=torch.uint8), pts.new_empty((N, 5))


# ------------------------------------------------------------------------------------------------ render_rays
@custom_op("vanerf_b200::render_rays", mutates_args=())
def render_rays(ctx: int, tar: torch.Tensor, pix_xy: torch.Tensor, ztab: torch.Tensor, utab: torch.Tensor, fine: bool,
                precision: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """VANeRF.batch_render_pifu_nerf's ray batch in one call (src/model.py:1103-1422, inference branch): rays -> coarse pass ->
    composite -> importance -> fine pass -> composite.  ztab (n_coarse), utab (n_fine) -> (R,8) rows [r,g,b,depth,alpha,sdf,0,0]
    of the coarse and the fine pass (the latter has 0 rows when fine is False)."""
    r = _renderer(ctx)
    R = pix_xy.shape[0]
    oc = r.empty((R, 8))
    of = r.empty((R, 8)) if fine else r.empty((0, 8))
    t = unpack_target(tar)
    st = r.lib.dll.vanerf_render_rays(r.ctx, precision, C.byref(t), _p(r, pix_xy), R, ztab.shape[0], utab.shape[0], int(fine), _p(r, ztab),
                                      _p(r, utab) if fine else None, _p(r, oc), _p(r, of) if fine else None, r.stream)
    _chk(r, st, "vanerf_render_rays")
    return oc, of


@render_rays.register_fake
def _(ctx, tar, pix_xy, ztab, utab, fine, precision):
    R = pix_xy.shape[0]
    return pix_xy.new_empty((R, 8), dtype=torch.float32), pix_xy.new_empty((R if fine else 0, 8), dtype=torch.float32)
