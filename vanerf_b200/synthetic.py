"""Synthetic two-hand scenes for the VANeRF render path.

InterHand2.6M, the MANO pickles and the released checkpoint are not available offline, so every
input of the path is generated here from a seed, with numpy only (bit-stable across machines):

* mesh      - MANO *topology* (2 x 779 vertices, 2 x 1554 faces = 778/1538 + wrist seal,
              reference `src/dataset.py:35-52`), closed genus-0 surface per hand.
* keypoints - 21 per hand (reference uses a joint regressor `kpt = J @ verts`; here fixed vertex
              group averages so that no reference asset has to travel).
* cameras   - on a 1 m sphere looking at the origin, pinhole K; layouts follow `decode_batch`
              (`src/model.py:306-356`).
* maps      - source images U(0,1), foreground masks, CNN feature maps N(0,1) with the shapes the
              reference encoders produce (`src/networks.py:76-78`).

Everything is returned as numpy float32 arrays; `to_torch` turns a scene into the dictionaries the
reference call surface takes.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

N_VERT_HAND = 779          # 778 MANO vertices + 1 wrist-seal vertex (src/networks.py:25)
N_FACE_HAND = 1554         # 1538 MANO faces + 16 seal faces
N_VERT = 2 * N_VERT_HAND   # 1558
N_FACE = 2 * N_FACE_HAND   # 3108
N_KPT = 42
ZNEAR, ZFAR = 0.71, 1.42   # src/model.py:58

SEED = 20240101


def _fibonacci_sphere(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    th = np.pi * (1.0 + 5.0 ** 0.5) * i
    return np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)


_HULL_CACHE: Dict[int, np.ndarray] = {}


def _sphere_faces(n: int) -> np.ndarray:
    """Outward-oriented triangulation of n Fibonacci points on the unit sphere (2n-4 faces)."""
    if n not in _HULL_CACHE:
        from scipy.spatial import ConvexHull
        p = _fibonacci_sphere(n)
        f = ConvexHull(p).simplices.astype(np.int64)
        c = p[f].mean(1)
        nrm = np.cross(p[f[:, 1]] - p[f[:, 0]], p[f[:, 2]] - p[f[:, 0]])
        flip = (nrm * c).sum(1) < 0
        f[flip] = f[flip][:, [0, 2, 1]]
        # canonical order (ConvexHull's simplex order is implementation defined)
        f = np.stack([np.roll(t, -int(np.argmin(t))) for t in f])
        order = np.lexsort((f[:, 2], f[:, 1], f[:, 0]))
        _HULL_CACHE[n] = f[order]
        assert _HULL_CACHE[n].shape[0] == 2 * n - 4
    return _HULL_CACHE[n]


def make_hand(center, radii, seed: int, bump: float = 0.15) -> Tuple[np.ndarray, np.ndarray]:
    """One closed 'hand': star-shaped deformation of an ellipsoid, 779 verts / 1554 faces."""
    rng = np.random.RandomState(seed)
    p = _fibonacci_sphere(N_VERT_HAND)
    # low-frequency radial displacement breaks convexity but keeps the surface star-shaped/closed
    fr = rng.uniform(1.0, 3.0, size=(3, 3))
    ph = rng.uniform(0, 2 * np.pi, size=(3,))
    r = 1.0 + bump * (np.sin(p @ fr[0] + ph[0]) * np.sin(p @ fr[1] + ph[1]) + 0.5 * np.sin(p @ fr[2] + ph[2]))
    v = p * r[:, None] * np.asarray(radii)[None] + np.asarray(center)[None]
    return v.astype(np.float32), _sphere_faces(N_VERT_HAND).copy()


def look_at(cam_center: np.ndarray, target=(0.0, 0.0, 0.0), up=(0.0, -1.0, 0.0)) -> np.ndarray:
    """World-to-camera [R|t] (3,4), OpenCV convention (+z forward, +y down)."""
    c = np.asarray(cam_center, np.float64)
    z = np.asarray(target, np.float64) - c
    z /= np.linalg.norm(z)
    x = np.cross(np.asarray(up, np.float64), z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z])
    t = -R @ c
    return np.concatenate([R, t[:, None]], 1).astype(np.float32)


def orbit_cam(az_deg: float, el_deg: float = 0.0, radius: float = 1.0) -> np.ndarray:
    az, el = math.radians(az_deg), math.radians(el_deg)
    c = radius * np.array([math.sin(az) * math.cos(el), -math.sin(el), -math.cos(az) * math.cos(el)])
    return look_at(c)


@dataclass
class Scene:
    H: int
    W: int
    V: int
    verts: np.ndarray            # (1558,3) f32 world
    faces: np.ndarray            # (3108,3) i64
    kpt3d: np.ndarray            # (42,3)
    bounds: np.ndarray           # (2,3)
    K_src: np.ndarray            # (V,3,3)
    Rt_src: np.ndarray           # (V,3,4)
    K_tar: np.ndarray            # (3,3)
    Rt_tar: np.ndarray           # (3,4)
    img: np.ndarray              # (V,3,H,W)
    fg_mask: np.ndarray          # (V,1,H,W) bool
    feat_geo0: np.ndarray        # (V,64,h0,w0)
    feat_geo1: np.ndarray        # (V,8,h1,w1)
    feat_tex: np.ndarray         # (V,8,h2,w2)
    meta: dict = field(default_factory=dict)


def feature_map_shapes(H: int, W: int) -> List[Tuple[int, int, int]]:
    """(C,h,w) of geo0, geo1, tex for an HxW source image (H/8 x ceil(W/8), ...)."""
    c8 = lambda a: -(-a // 8)
    c4 = lambda a: -(-a // 4)
    c2 = lambda a: -(-a // 2)
    return [(64, c8(H), c8(W)), (8, c2(H), c2(W)), (8, c4(H), c4(W))]


def _silhouette(verts, K, Rt, H, W, dilate=6) -> np.ndarray:
    """Cheap foreground mask: splat projected vertices and dilate."""
    p = verts @ Rt[:, :3].T + Rt[:, 3]
    uv = (p @ K.T)
    uv = uv[:, :2] / uv[:, 2:3]
    m = np.zeros((H, W), bool)
    x = np.clip(np.round(uv[:, 0]).astype(int), 0, W - 1)
    y = np.clip(np.round(uv[:, 1]).astype(int), 0, H - 1)
    m[y, x] = True
    # separable box dilation
    acc = m.copy()
    for d in range(1, dilate + 1):
        acc[:, d:] |= m[:, :-d]
        acc[:, :-d] |= m[:, d:]
    m2 = acc.copy()
    for d in range(1, dilate + 1):
        m2[d:, :] |= acc[:-d, :]
        m2[:-d, :] |= acc[d:, :]
    return m2


def make_scene(H: int = 512, W: int = 334, V: int = 3, seed: int = SEED, layout: str = "narrow",
               frame: int = 0, tar_az: float = 0.0, tar_el: float = 0.0, mask: str = "silhouette",
               focal: float | None = None) -> Scene:
    """layout 'narrow' = vanerf.json-like small baselines (+-15, 30 deg); 'bvv' = big view variation
    (+-60, 180 deg), hands overlapping in depth from the target (configs/vanerf_bvv.json)."""
    rng = np.random.RandomState(seed + 7919 * frame)
    if layout == "bvv":
        centers = [np.array([0.030, 0.0, -0.035]), np.array([-0.030, 0.0, 0.035])]
        az_src = [60.0, -60.0, 180.0, 120.0][:V]
    else:
        centers = [np.array([0.045, 0.0, 0.0]), np.array([-0.045, 0.0, 0.0])]
        az_src = [15.0, -15.0, 30.0, -30.0][:V]
    radii = np.array([0.05, 0.09, 0.03])
    vr, fr = make_hand(centers[0], radii, seed)
    vl, fl = make_hand(centers[1], radii, seed + 1)
    verts = np.concatenate([vr, vl]).astype(np.float32)
    if frame:
        # smooth per-frame displacement, amplitude 1 cm (config D, render_dynamic)
        ph = 0.37 * frame
        disp = 0.01 * np.stack([np.sin(20 * verts[:, 1] + ph), np.sin(15 * verts[:, 2] + 1.3 * ph),
                                np.sin(25 * verts[:, 0] + 0.7 * ph)], 1)
        verts = (verts + disp).astype(np.float32)
    faces = np.concatenate([fr, fl + N_VERT_HAND]).astype(np.int64)
    # keypoints: 21 per hand, each the mean of 8 fixed vertices (stands in for J_regressor @ verts)
    krng = np.random.RandomState(1234)
    groups = krng.randint(0, N_VERT_HAND - 1, size=(21, 8))
    kpt = np.concatenate([verts[:N_VERT_HAND][groups].mean(1), verts[N_VERT_HAND:][groups].mean(1)]).astype(np.float32)
    mn, mx = verts.min(0).copy(), verts.max(0).copy()
    mn[2] -= 0.05
    mx[2] += 0.05                                  # src/dataset.py:191-195
    bounds = np.stack([mn, mx]).astype(np.float32)
    if focal is None:
        focal = 1100.0 if max(H, W) > 256 else 700.0
    K = np.array([[focal, 0, W / 2.0], [0, focal, H / 2.0], [0, 0, 1]], np.float32)
    Rt_src = np.stack([orbit_cam(a) for a in az_src]).astype(np.float32)
    Rt_tar = orbit_cam(tar_az, tar_el)
    img = rng.uniform(0, 1, size=(V, 3, H, W)).astype(np.float32)
    if mask == "ones":
        fg = np.ones((V, 1, H, W), bool)
    else:
        fg = np.stack([_silhouette(verts, K, Rt_src[v], H, W)[None] for v in range(V)])
    shp = feature_map_shapes(H, W)
    g0 = rng.standard_normal((V,) + shp[0]).astype(np.float32)
    g1 = rng.standard_normal((V,) + shp[1]).astype(np.float32)
    tx = rng.standard_normal((V,) + shp[2]).astype(np.float32)
    return Scene(H=H, W=W, V=V, verts=verts, faces=faces, kpt3d=kpt, bounds=bounds,
                 K_src=np.repeat(K[None], V, 0), Rt_src=Rt_src, K_tar=K.copy(), Rt_tar=Rt_tar,
                 img=img, fg_mask=fg, feat_geo0=g0, feat_geo1=g1, feat_tex=tx,
                 meta=dict(seed=seed, layout=layout, frame=frame, focal=float(focal)))


def to_torch(scene: Scene, device="cpu"):
    """Build the reference-layout dictionaries (`decode_batch`, src/model.py:306-378; B = 1)."""
    import torch
    V, H, W = scene.V, scene.H, scene.W
    t = lambda a, dt=torch.float32: torch.from_numpy(np.ascontiguousarray(a)).to(device=device, dtype=dt)
    extrin = torch.eye(4, device=device)[None].repeat(V, 1, 1)
    extrin[:, :3, :4] = t(scene.Rt_src)
    intrin = torch.eye(4, device=device)[None].repeat(V, 1, 1)
    intrin[:, :3, :3] = t(scene.K_src)
    cam_in = {"KRT": torch.bmm(intrin, extrin), "K": intrin, "Rt": t(scene.Rt_src), "extrin": extrin,
              "znear": ZNEAR, "zfar": ZFAR, "width": W, "height": H, "nml_scale": 100.0}
    e_t = torch.eye(4, device=device)[None].clone()
    e_t[:, :3, :4] = t(scene.Rt_tar)[None]
    i_t = torch.eye(4, device=device)[None].clone()
    i_t[:, :3, :3] = t(scene.K_tar)[None]
    cam_tar = {"K": i_t, "RT": e_t, "KRT": torch.bmm(i_t, e_t), "width": W, "height": H,
               "nml_scale": 100.0, "znear": ZNEAR, "zfar": ZFAR}
    targets = {"vert_world": t(scene.verts)[None], "face_world": t(scene.faces.astype(np.float32))[None],
               "tar_cam": {"tar_R": torch.eye(3, device=device)[None], "tar_T": torch.zeros(1, 3, device=device),
                           "tar_focal": torch.ones(1, 2, device=device), "tar_princpt": torch.ones(1, 2, device=device)}}
    sp_data = {"extrin": extrin, "kpt3d": t(scene.kpt3d)[None]}
    return dict(
        img=t(scene.img), cam_in=cam_in, cam_tar=cam_tar, targets=targets, sp_data=sp_data,
        hand_type=torch.ones(1, 2, device=device),
        feat_geo=[t(scene.feat_geo0), t(scene.feat_geo1)], feat_tex=t(scene.feat_tex),
        src_foreground_mask=t(scene.fg_mask, torch.bool)[None], bounds=t(scene.bounds)[None],
        objcenter=t(scene.kpt3d)[None][:, 0],
    )
