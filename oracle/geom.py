"""ORACLE — test infrastructure only (see oracle/README.md).  ctypes front-end of geom_oracle.c.

Torch-facing stand-ins with the call signatures of the kaolin / pytorch3d entry points that the
reference calls (src/lib/dataset/mesh_util.py:498-524, src/networks.py:27-33).  PARITY UNPINNED: the
libraries are not installable here; semantics are defined in geom_oracle.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgeom_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "geom_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        P, I64 = ctypes.c_void_p, ctypes.c_int64
        _lib.vo_point_mesh_distance.argtypes = [P, I64, P, P, I64, P, P]
        _lib.vo_check_sign.argtypes = [P, I64, P, P, I64, P]
        _lib.vo_knn1.argtypes = [P, I64, P, I64, P]
        _lib.vo_rasterize.argtypes = [P, P, I64, ctypes.c_int, ctypes.c_int, ctypes.c_int, P]
        for f in (_lib.vo_point_mesh_distance, _lib.vo_check_sign, _lib.vo_knn1, _lib.vo_rasterize):
            f.restype = None
    return _lib


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _i64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int64))


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _split(n, parts):
    step = -(-n // parts)
    return [(s, min(n, s + step)) for s in range(0, n, step)]


_NT = max(1, os.cpu_count() or 1)


def point_mesh_distance_np(pts, verts, faces):
    pts, verts, faces = _f32(pts).reshape(-1, 3), _f32(verts).reshape(-1, 3), _i64(faces).reshape(-1, 3)
    n = pts.shape[0]
    d2 = np.empty(n, np.float32)
    idx = np.empty(n, np.int64)
    L = lib()

    def run(se):
        s, e = se
        L.vo_point_mesh_distance(_ptr(pts[s:e]), e - s, _ptr(verts), _ptr(faces), faces.shape[0],
                                 _ptr(d2[s:e]), _ptr(idx[s:e]))
    with ThreadPoolExecutor(_NT) as ex:
        list(ex.map(run, _split(n, _NT * 4)))
    return d2, idx


def check_sign_np(pts, verts, faces):
    pts, verts, faces = _f32(pts).reshape(-1, 3), _f32(verts).reshape(-1, 3), _i64(faces).reshape(-1, 3)
    n = pts.shape[0]
    out = np.empty(n, np.uint8)
    L = lib()

    def run(se):
        s, e = se
        L.vo_check_sign(_ptr(pts[s:e]), e - s, _ptr(verts), _ptr(faces), faces.shape[0], _ptr(out[s:e]))
    with ThreadPoolExecutor(_NT) as ex:
        list(ex.map(run, _split(n, _NT * 4)))
    return out.astype(bool)


def knn1_np(pts, verts):
    pts, verts = _f32(pts).reshape(-1, 3), _f32(verts).reshape(-1, 3)
    n = pts.shape[0]
    idx = np.empty(n, np.int64)
    L = lib()

    def run(se):
        s, e = se
        L.vo_knn1(_ptr(pts[s:e]), e - s, _ptr(verts), verts.shape[0], _ptr(idx[s:e]))
    with ThreadPoolExecutor(_NT) as ex:
        list(ex.map(run, _split(n, _NT * 4)))
    return idx


def rasterize_np(xyz, faces, image_size=256):
    xyz, faces = _f32(xyz).reshape(-1, 3), _i64(faces).reshape(-1, 3)
    out = np.empty((image_size, image_size), np.int64)
    L = lib()

    def run(se):
        L.vo_rasterize(_ptr(xyz), _ptr(faces), faces.shape[0], int(image_size), se[0], se[1], _ptr(out))
    with ThreadPoolExecutor(_NT) as ex:
        list(ex.map(run, _split(image_size, _NT * 4)))
    return out


# ---- torch-facing stand-ins with the third-party signatures ------------------------------------------------

def index_vertices_by_faces(verts, faces):          # kaolin.ops.mesh.index_vertices_by_faces
    return verts[:, faces.long()]


def point_to_mesh_distance(points, face_vertices):  # kaolin.metrics.trianglemesh.point_to_mesh_distance
    ds, is_ = [], []
    for b in range(points.shape[0]):
        fv = face_vertices[b].detach().cpu().numpy().astype(np.float32)      # (F,3,3)
        verts = fv.reshape(-1, 3)
        faces = np.arange(verts.shape[0], dtype=np.int64).reshape(-1, 3)
        d2, idx = point_mesh_distance_np(points[b].detach().cpu().numpy(), verts, faces)
        ds.append(torch.from_numpy(d2))
        is_.append(torch.from_numpy(idx))
    d, i = torch.stack(ds).to(points.device), torch.stack(is_).to(points.device)
    return d, i, torch.zeros_like(i, dtype=torch.int32)


def check_sign(verts, faces, points):               # kaolin.ops.mesh.check_sign
    res = []
    for b in range(points.shape[0]):
        res.append(torch.from_numpy(check_sign_np(points[b].detach().cpu().numpy(),
                                                  verts[b].detach().cpu().numpy(), faces.detach().cpu().numpy())))
    return torch.stack(res).to(points.device)


def knn_points(p1, p2, K=1, **kw):                   # pytorch3d.ops.knn_points
    assert K == 1
    idx = []
    for b in range(p1.shape[0]):
        idx.append(torch.from_numpy(knn1_np(p1[b].detach().cpu().numpy(), p2[b].detach().cpu().numpy())))
    idx = torch.stack(idx)[..., None].to(p1.device)
    d = ((p1 - torch.gather(p2, 1, idx.expand(-1, -1, 3))) ** 2).sum(-1, keepdim=True)
    return d, idx, None


class Meshes:                                        # pytorch3d.structures.Meshes (only what get_visibility uses)
    def __init__(self, verts, faces, textures=None):
        self.v, self.f = verts, faces

    def to(self, d):
        return self


def rasterize_meshes(meshes, image_size=256, blur_radius=0.0, faces_per_pixel=1, bin_size=None,
                     max_faces_per_bin=None, perspective_correct=False, cull_backfaces=False, **kw):
    assert faces_per_pixel == 1 and blur_radius == 0.0 and perspective_correct and cull_backfaces
    p2f = rasterize_np(meshes.v[0].detach().cpu().numpy(), meshes.f[0].detach().cpu().numpy(), image_size)
    p2f = torch.from_numpy(p2f).view(1, image_size, image_size, 1)
    return p2f, None, None, None
