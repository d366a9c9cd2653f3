"""ORACLE — test infrastructure only.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; nothing under vanerf_b200/ does.

Self-contained CPU restatement (numpy + torch-CPU, fp32) of VANeRF's novel-view render path, V-generalised
 and free of the 256x256 assumptions, following SURVEY.md Appendix A step by step:

  rays / box clip / coarse samples   src/model.py:1190-1238, :1497-1570
  vertex projection + visibility     src/model.py:1244-1255, src/lib/dataset/mesh_util.py:284-318,484-489
  signed distance / query visibility src/lib/dataset/mesh_util.py:498-524, :321-356
  projection, masks, pix_weight      src/model.py:768-821
  feat_sample                        src/utils.py:136-151
  SpatialEncoder rel_z_decay         src/spatial.py:59-117
  KNN_vis / GeoVisFusion             src/networks.py:27-33, :75-106
  MLPUNetFusion                      src/utils.py:633-649, :822-880
  query_color / TexVisFusion / IBR   src/model.py:884-957, src/networks.py:268-293, src/model.py:1600-1636
  eval_func / sdf_activation         src/model.py:1140-1160, :879-882
  rgba2out / importance_sample       src/model.py:1465-1494, :1425-1462

Pinned against the reference itself: tests/test_oracle_vs_reference.py (build container, reference imported
through oracle/ref_import.py) and the committed golden vectors tests/golden/*.npz (generated from the
reference by tests/golden/make_golden.py).  The geometry queries behind kaolin / pytorch3d are PARITY
UNPINNED (see oracle/geom_oracle.c).

Arithmetic contract for the bit-exact quantities (sample depths/positions, masks, indices): every step is a
single correctly-rounded fp32 operation in the order written here, no FMA contraction, 3-term dot products
as (a0*b0 + a1*b1) + a2*b2; the only fused operations are the bilinear tap accumulation of `grid_sample`,
which torch-CPU evaluates as fma(se,SE, fma(sw,SW, fma(ne,NE, nw*NW))), and 3-vector 2-norms, which it
evaluates as sqrt(fma(z,z, fma(y,y, x*x))) (both verified bit-for-bit against torch 2.11 CPU).
Per-frame 3x3 / 4x4 matrices (inverse(K), KRT, -t^T R) come from the same torch calls the reference makes.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import geom

f32 = np.float32
NUM_V = 779            # src/networks.py:25


# ------------------------------------------------------------------------------------------------ helpers
def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def _n(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def dot3(a0, a1, a2, b0, b1, b2):
    return (a0 * b0 + a1 * b1) + a2 * b2


def fma_np(a, b, c):
    """fp32 fused multiply-add emulated through fp64 (product exact, one extra rounding at 2^-53)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def norm3(x, y, z):
    """torch-CPU 2-norm of a 3-vector: sqrt(fma(z,z, fma(y,y, x*x))) (verified bit-for-bit on vector_norm)."""
    return np.sqrt(fma_np(z, z, fma_np(y, y, (x * x).astype(f32))))


def affine_rows(p, M):
    """p (N,3) f32, M (3or4,4) -> (N,3): ((p0*m0 + p1*m1) + p2*m2) + m3  (v @ M[:3,:3].T + M[:3,3])."""
    p0, p1, p2 = p[:, 0], p[:, 1], p[:, 2]
    return np.stack([dot3(p0, p1, p2, M[j, 0], M[j, 1], M[j, 2]) + M[j, 3] for j in range(3)], 1).astype(f32)


def bilinear_np(feat, xy):
    """grid_sample(bilinear, border, align_corners=True) restated.  feat (C,Hf,Wf) f32, xy (N,2) in [-1,1] -> (N,C)."""
    feat = np.asarray(feat, f32)
    C, Hf, Wf = feat.shape
    one = f32(1)
    x, y = xy[:, 0].astype(f32), xy[:, 1].astype(f32)
    ix = ((x + one) / f32(2)) * f32(Wf - 1)
    iy = ((y + one) / f32(2)) * f32(Hf - 1)
    # clip_coordinates: min(size-1, max(x, 0)); NaN -> 0 like std::max/min on the vectorised path is not
    # relied upon (coordinates on this path are finite)
    ix = np.minimum(f32(Wf - 1), np.maximum(ix, f32(0)))
    iy = np.minimum(f32(Hf - 1), np.maximum(iy, f32(0)))
    x0, y0 = np.floor(ix), np.floor(iy)
    w = ix - x0
    e = one - w
    n_ = iy - y0
    s = one - n_
    NW, NE, SW, SE = s * e, s * w, n_ * e, n_ * w
    x0i, y0i = x0.astype(np.int64), y0.astype(np.int64)
    x1i, y1i = x0i + 1, y0i + 1

    def tap(yy, xx):
        ok = (xx < Wf) & (yy < Hf)
        v = feat[:, np.minimum(yy, Hf - 1), np.minimum(xx, Wf - 1)].T
        return np.where(ok[:, None], v, f32(0)).astype(f32)
    a, b, c, d = tap(y0i, x0i), tap(y0i, x1i), tap(y1i, x0i), tap(y1i, x1i)
    bc = lambda t: np.broadcast_to(t[:, None], a.shape)
    return fma_np(d, bc(SE), fma_np(c, bc(SW), fma_np(b, bc(NE), (a * bc(NW)).astype(f32))))


def softplus100(x):
    return F.softplus(x, beta=100, threshold=20)


# ------------------------------------------------------------------------------------------------ rays
def pixel_grid(H, W, level, stride_xy=(0, 0)):
    """src/model.py:1191-1200 (inference branch).  Returns int64 (R,2) [x,y] and the flat index."""
    step = 2 ** (level - 1)
    ys, xs = np.meshgrid(np.arange(0, H, step), np.arange(0, W, step), indexing="ij")
    g = np.stack([xs, ys], -1).reshape(-1, 2).astype(np.int64) + np.asarray(stride_xy, np.int64)[None]
    return g, g[:, 0] + g[:, 1] * W


def frame_camera(cam_tar):
    """Per-frame matrices by the same torch calls as src/model.py:1208,1213."""
    K, RT = cam_tar["K"].float().cpu(), cam_tar["RT"].float().cpu()
    inv_K = torch.inverse(K[:, :3, :3]).transpose(1, 2)[0].numpy().astype(f32)      # (3,3): d = g @ inv_K
    R = RT[0, :3, :3].numpy().astype(f32)
    cam_pos = (-torch.bmm(RT[:, :3, 3][:, None], RT[:, :3, :3]))[0, 0].numpy().astype(f32)
    return inv_K, R, cam_pos


def make_rays(grids_xy, inv_K, R, cam_pos, znear, zfar, bounds):
    """A.1 steps 2-3.  grids_xy (R,2) ints.  Returns dict of fp32 arrays."""
    x = grids_xy[:, 0].astype(f32)
    y = grids_xy[:, 1].astype(f32)
    one = np.ones_like(x)
    zn, zf = f32(znear), f32(zfar)

    def cam_dir(gx, gy, gz):
        return [dot3(gx, gy, gz, inv_K[0, j], inv_K[1, j], inv_K[2, j]) for j in range(3)]
    d = cam_dir(x, y, one)
    dn = cam_dir(zn * x, zn * y, zn * one)
    df = cam_dir(zf * x, zf * y, zf * one)
    znear_rays = norm3(dn[0], dn[1], dn[2])
    zfar_rays = norm3(df[0], df[1], df[2])
    w = [dot3(d[0], d[1], d[2], R[0, j], R[1, j], R[2, j]) for j in range(3)]
    nrm = norm3(w[0], w[1], w[2])
    nrm = np.maximum(nrm, f32(1e-12))
    dirs = np.stack([w[0] / nrm, w[1] / nrm, w[2] / nrm], 1).astype(f32)

    # --- ray_bbox_intersection (src/model.py:1497-1570), boffset (-0.01, 0.01)
    b = np.asarray(bounds, f32).reshape(2, 3) + np.asarray([[-0.01], [0.01]], f32)
    dd = dirs.copy()
    dd[np.abs(dd) < f32(1e-5)] = f32(1e-5)
    o = cam_pos.astype(f32)
    t6 = np.concatenate([(b[0][None] - o[None]) / dd, (b[1][None] - o[None]) / dd], 1).astype(f32)   # (R,6)
    p6 = t6[:, :, None] * dd[:, None, :] + o[None, None, :]                                          # (R,6,3)
    eps = f32(1e-6)
    lo, hi = b[0] - eps, b[1] + eps
    inside = np.ones(p6.shape[:2], bool)
    for c in range(3):
        inside &= (p6[:, :, c] >= lo[c]) & (p6[:, :, c] <= hi[c])
    hit = inside.sum(1) == 2
    nr = norm3(dd[:, 0], dd[:, 1], dd[:, 2])
    near = np.ones(len(x), f32)
    far = np.ones(len(x), f32)
    if hit.any():
        idx = np.argsort(~inside[hit], axis=1, kind="stable")[:, :2]        # the two inside faces, in 6-order
        pts = np.take_along_axis(p6[hit], idx[:, :, None], 1)               # (h,2,3)
        dl = pts - o[None, None, :]
        dist = norm3(dl[..., 0], dl[..., 1], dl[..., 2]) / nr[hit][:, None]
        near[hit] = dist.min(1)
        far[hit] = dist.max(1)
    m1 = (hit & (near > znear_rays)).astype(f32)
    znear_c = m1 * near + (one - m1) * znear_rays
    m2 = (hit & (far < zfar_rays)).astype(f32)
    zfar_c = m2 * far + (one - m2) * zfar_rays
    return dict(dirs=dirs, znear=znear_c.astype(f32), zfar=zfar_c.astype(f32), hit=hit,
                box_near=near, box_far=far, frustum_near=znear_rays, frustum_far=zfar_rays)


def sample_z(znear_rays, zfar_rays, t_tab):
    """A.1-4 inference branch: z = znear + (zfar - znear) * t."""
    return (znear_rays[:, None] + (zfar_rays - znear_rays)[:, None] * t_tab[None, :]).astype(f32)


def sample_points(cam_pos, dirs, z):
    """eval_pts = cam_pos + dir * z (mul, then add); sample index fastest.  -> (R*S,3)."""
    p = cam_pos[None, None, :] + dirs[:, None, :] * z[:, :, None]
    return p.reshape(-1, 3).astype(f32)


# ------------------------------------------------------------------------------------------------ per-frame geometry
def project_vertices(verts, KRT, W, H, znear, zfar):
    """A.2: normalised [0,1] image coords + depth used by the visibility raster; and [-1,1] coords used for
    vertex-feature sampling (src/model.py:845-853)."""
    vimg = affine_rows(verts, KRT)
    z = vimg[:, 2]
    vx = vimg[:, 0] / (z + f32(1e-8))
    vy = vimg[:, 1] / (z + f32(1e-8))
    xy01 = np.stack([vx / f32(W - 1.0), vy / f32(H - 1.0)], 1).astype(f32)
    z01 = ((z - f32(znear)) / f32(zfar - znear)).astype(f32)
    xy11 = np.stack([f32(2.0) * (vx / f32(W - 1.0)) - f32(1.0), f32(2.0) * (vy / f32(H - 1.0)) - f32(1.0)], 1).astype(f32)
    return xy01, z01, xy11


def vertex_visibility(xy01, z01, faces):
    """get_visibility (mesh_util.py:284-318): raster at 256, visible faces -> visible vertices; the -1 entry of
    unique(pix_to_face) indexes the LAST face."""
    xyz = ((np.concatenate([xy01, z01[:, None]], 1) + f32(1.0)) / f32(2.0)).astype(f32)
    p2f = geom.rasterize_np(xyz, faces, 256)
    uf = np.unique(p2f)
    vis_v = np.unique(faces[uf])            # negative index wraps to the last face, like torch
    vis = np.zeros(xy01.shape[0], f32)
    vis[vis_v] = 1.0
    return vis, p2f


def signed_distance(pts, verts, faces):
    d2, fidx = geom.point_mesh_distance_np(pts, verts, faces)
    dist = np.sqrt(d2 + f32(1e-6))
    inside = geom.check_sign_np(pts, verts, faces)
    sign = f32(-2.0) * (inside.astype(f32) - f32(0.5))
    return (dist * sign).astype(f32), fidx, inside


def query_visibility(pts, verts, faces, fidx, vert_vis):
    """barycentric_coordinates_of_projection (mesh_util.py:321-356) + blend >= 0.1 (mesh_util.py:515-522)."""
    tri = verts[faces[fidx]]                 # (N,3,3)
    v0, v1, v2 = tri[:, 0], tri[:, 1], tri[:, 2]
    u, v = v1 - v0, v2 - v0

    def cross(a, b):
        return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                         a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1).astype(f32)

    def sum3(a):
        return (a[:, 0] + a[:, 1]) + a[:, 2]
    n = cross(u, v)
    s = sum3(n * n)
    s = np.where(s == 0, f32(1e-6), s).astype(f32)
    inv = f32(1.0) / s
    w = pts - v0
    b2 = sum3(cross(u, w) * n) * inv
    b1 = sum3(cross(w, v) * n) * inv
    b0 = (f32(1.0) - b1) - b2
    cv = vert_vis[faces[fidx]]               # (N,3)
    blend = (cv[:, 0] * b0 + cv[:, 1] * b1) + cv[:, 2] * b2
    return blend >= f32(1e-1)


# ------------------------------------------------------------------------------------------------ network
class OracleNet:
    """Holds folded weights (vanerf_b200.weights.fold layout is NOT imported: folding is restated here)."""

    def __init__(self, state_dict: Dict[str, np.ndarray]):
        sd = {k[6:] if k.startswith("model.") else k: _t(np.asarray(_n(v), f32)) for k, v in state_dict.items()}
        self.sd = sd

        def wn(p):
            v, g = sd[p + "weight_v"], sd[p + "weight_g"]
            return torch._weight_norm(v, g, 0)
        c1 = lambda k: sd[k][:, :, 0]
        self.geo = [dict(at0=c1("geo_vis_fusion.fconv_at.0.weight"), at1=c1("geo_vis_fusion.fconv_at.2.weight"),
                         f0=c1("geo_vis_fusion.fconv_ated.0.weight"), f1=c1("geo_vis_fusion.fconv_ated.2.weight")),
                    dict(at0=c1("geo_vis_fusion.fconv_at1.0.weight"), at1=c1("geo_vis_fusion.fconv_at1.2.weight"),
                         f0=c1("geo_vis_fusion.fconv_ated1.0.weight"), f1=c1("geo_vis_fusion.fconv_ated1.2.weight"))]
        self.mlp = []
        for i in range(3):
            p = f"mlp_geo.layers1.layers.{i}.linear."
            self.mlp.append((wn(p), sd[p + "bias"]))
        self.mlp.append((sd["mlp_geo.layers1.layers.3.linear.weight"], sd["mlp_geo.layers1.layers.3.linear.bias"]))
        self.post = []
        for i in range(2):
            p = f"mlp_geo.layers2.layers.{i}.linear."
            self.post.append((wn(p), sd[p + "bias"]))
        self.post.append((sd["mlp_geo.layers2.layers.2.linear.weight"], sd["mlp_geo.layers2.layers.2.linear.bias"]))
        self.compress = (sd["ibr_compress_gfeat.weight"], sd["ibr_compress_gfeat.bias"])
        self.tex = dict(f0=c1("tex_vis_fusion.fconv.0.weight"), f1=c1("tex_vis_fusion.fconv.2.weight"),
                        at0=c1("tex_vis_fusion.fconv_at.0.weight"), at1=c1("tex_vis_fusion.fconv_at.2.weight"))
        g = lambda n, j: (sd[f"mlp_tex.{n}.{j}.weight"], sd[f"mlp_tex.{n}.{j}.bias"])
        self.ibr = dict(ray=[g("ray_encoder", 0), g("ray_encoder", 2)], base=[g("base_layer", 0), g("base_layer", 2)],
                        vis1=[g("vis_layer1", 0), g("vis_layer1", 2)], vis2=[g("vis_layer2", 0), g("vis_layer2", 2)],
                        out=[g("out_layer", 0), g("out_layer", 2), g("out_layer", 4)], ani_al=sd["mlp_tex.ani_al"])
        self.sigmoid_beta = float(max(2e-3, float(sd["sigmoid_beta"].reshape(-1)[0])))     # model.py:880

    # ---- per-frame: TexVisFusion global feature (src/networks.py:273-279), torch ops like the reference
    def global_vertex_feature(self, img, feat_tex):
        sd = self.sd

        def stack(x, pre):
            x = F.conv2d(x, sd[pre + ".0.weight"], padding=1)
            x = F.relu(F.layer_norm(x, x.shape[-2:], sd[pre + ".1.weight"], sd[pre + ".1.bias"], 1e-6))
            x = F.conv2d(x, sd[pre + ".3.weight"], padding=1)
            x = F.relu(F.layer_norm(x, x.shape[-2:], sd[pre + ".4.weight"], sd[pre + ".4.bias"], 1e-6))
            return F.adaptive_avg_pool2d(x, 3)
        gf = stack(feat_tex, "tex_vis_fusion.fconv3")
        gf = gf.reshape(*gf.shape[:2], -1)
        gi = stack(img, "tex_vis_fusion.fconv4")
        gi = gi.reshape(*gi.shape[:2], -1)
        g = torch.cat([gi, gf], -1)                                            # (V,42,18)
        p = "tex_vis_fusion.fconv_gt"
        x = F.conv1d(g, sd[p + ".0.weight"], padding=1)
        x = F.relu(F.layer_norm(x, (18,), sd[p + ".1.weight"], sd[p + ".1.bias"], 1e-6))
        x = F.conv1d(x, sd[p + ".3.weight"], padding=1)
        x = F.relu(F.layer_norm(x, (18,), sd[p + ".4.weight"], sd[p + ".4.bias"], 1e-6))
        return x                                                               # (V,1558,18)


class Oracle:
    """End-to-end restatement.  Inputs use the reference layouts (vanerf_b200.synthetic.to_torch)."""

    def __init__(self, state_dict, inp: dict):
        self.net = OracleNet(state_dict)
        self.inp = inp
        cam = inp["cam_in"]
        self.V = cam["KRT"].shape[0]
        self.W, self.H = int(cam["width"]), int(cam["height"])
        self.znear, self.zfar = float(cam["znear"]), float(cam["zfar"])
        self.KRT = _n(cam["KRT"]).astype(f32)
        self.extrin = _n(inp["sp_data"]["extrin"]).astype(f32)
        self.kpt3d = _n(inp["sp_data"]["kpt3d"])[0].astype(f32)
        self.verts = _n(inp["targets"]["vert_world"])[0].astype(f32)
        self.faces = _n(inp["targets"]["face_world"])[0].astype(np.int64)
        self.img = _n(inp["img"]).astype(f32)
        self.fg = _n(inp["src_foreground_mask"]).reshape(self.V, 1, self.H, self.W).astype(f32)
        self.geo0, self.geo1 = [_n(t).astype(f32) for t in inp["feat_geo"]]
        self.tex = _n(inp["feat_tex"]).astype(f32)
        self.frame = None

    # ---------------------------------------------------------------- per-frame state (A.2, A.3, A.6, A.8 tables)
    def frame_setup(self):
        V = self.V
        vis, xy11s, p2fs = [], [], []
        for v in range(V):
            xy01, z01, xy11 = project_vertices(self.verts, self.KRT[v], self.W, self.H, self.znear, self.zfar)
            vv, p2f = vertex_visibility(xy01, z01, self.faces)
            vis.append(vv)
            xy11s.append(xy11)
            p2fs.append(p2f)
        vert_vis = np.stack(vis)                                                         # (V,Nv)
        gfeat = _n(self.net.global_vertex_feature(_t(self.img), _t(self.tex)))           # (V,Nv,18)
        T64 = np.stack([bilinear_np(self.geo0[v], xy11s[v]) for v in range(V)])          # (V,Nv,64)
        T8 = np.stack([bilinear_np(self.geo1[v], xy11s[v]) for v in range(V)])           # (V,Nv,8)
        Ttex = np.stack([np.concatenate([bilinear_np(self.img[v], xy11s[v]), bilinear_np(self.tex[v], xy11s[v]),
                                         gfeat[v]], 1) for v in range(V)])               # (V,Nv,29)
        inv = torch.inverse(_t(self.KRT).float())                                        # model.py:937
        src_pos = inv[:, :3, 3].numpy().astype(f32)                                      # (V,3)
        self.frame = dict(vert_vis=vert_vis, vert_xy11=np.stack(xy11s), pix_to_face=np.stack(p2fs),
                          T64=T64, T8=T8, Ttex=Ttex, gfeat=gfeat, src_pos=src_pos)
        return self.frame

    # ---------------------------------------------------------------- per-sample geometry (A.3)
    def geometry(self, pts):
        fr = self.frame or self.frame_setup()
        sdf, fidx, inside = signed_distance(pts, self.verts, self.faces)
        nn = geom.knn1_np(pts, self.verts)
        qvis = np.stack([query_visibility(pts, self.verts, self.faces, fidx, fr["vert_vis"][v]) for v in range(self.V)])
        return dict(sdf=sdf, face=fidx, inside=inside, nn=nn, qvis=qvis)

    # ---------------------------------------------------------------- VANeRF.query (A.4-A.8)
    def query(self, pts, view, geo, n_pts_samples, taps: Optional[dict] = None):
        """pts (N,3), view (N,3) ray dirs, geo = self.geometry(pts).  Returns out (N,5)=[o0,o1,r,g,b], valid (N,)."""
        fr = self.frame or self.frame_setup()
        net, V, N = self.net, self.V, pts.shape[0]
        W, H = self.W, self.H
        one = f32(1.0)
        xy_all, zn_all, in_all, fg_all = [], [], [], []
        for v in range(V):
            vh = affine_rows(pts, self.KRT[v])
            z = vh[:, 2]
            x = vh[:, 0] / z
            y = vh[:, 1] / z
            x = f32(2.0) * (x / f32(W - 1.0)) - one
            y = f32(2.0) * (y / f32(H - 1.0)) - one
            zn = (f32(2.0) * (z - f32(self.znear))) / f32(self.zfar - self.znear) - one
            lo, hi = f32(-1.0 - 1e-2), f32(1.0 + 1e-2)
            inm = (x >= lo) & (x <= hi) & (y >= lo) & (y <= hi) & (zn >= f32(-1.0))
            xy = np.stack([x, y], 1).astype(f32)
            fgv = bilinear_np(self.fg[v], xy)[:, 0] > f32(0.1)
            xy_all.append(xy), zn_all.append(zn.astype(f32)), in_all.append(inm), fg_all.append(fgv)
        in_all, fg_all = np.stack(in_all), np.stack(fg_all)
        m = in_all.all(0) & fg_all.all(0)                                  # out_mask is identical for all views (B-10)
        out_mask = np.broadcast_to(m[None].astype(f32), (V, N)).copy()
        # pix_weight (model.py:813-821)
        pw = []
        for v in range(V):
            q = f32(0.5) * np.concatenate([xy_all[v], zn_all[v][:, None]], 1) + f32(0.5)
            d = np.minimum(q, one - q)
            s = 1.0 / (1.0 + np.exp(-(f32(5.0) * (d / f32(0.1) - one)).astype(np.float64)))
            s = s.astype(f32)
            pw.append((s[:, 0] * s[:, 1] * s[:, 2]) * out_mask[v])
        pw = np.stack(pw).astype(f32)
        pw = pw / (pw.sum(0, keepdims=True) + f32(1e-6))
        valid = out_mask.sum(0) > 0

        nn, twin = geo["nn"], (geo["nn"] + NUM_V) % (2 * NUM_V)
        sdf = _t(geo["sdf"])[:, None]
        lat_views, g64_all, g8_all = [], [], []
        pe_all = []
        for v in range(V):
            vis = fr["vert_vis"][v]
            qv = _t(geo["qvis"][v].astype(f32))[:, None]
            vn, vt = _t(vis[nn])[:, None], _t(vis[twin])[:, None]
            fused = []
            for c, (fmap, tab) in enumerate([(self.geo0[v], fr["T64"][v]), (self.geo1[v], fr["T8"][v])]):
                px = _t(bilinear_np(fmap, xy_all[v]))
                a = _t(tab[nn]) * vn
                b = _t(tab[twin]) * vt
                w = net.geo[c]
                x = torch.cat([px, a, b, sdf, qv, vn, vt], 1)
                att = torch.sigmoid(F.linear(F.relu(F.linear(x, w["at0"])), w["at1"]))
                x2 = torch.cat([px * att[:, 0:1], a * att[:, 1:2], b * att[:, 2:3], sdf, qv, vn, vt], 1)
                fused.append(F.linear(F.relu(F.linear(x2, w["f0"])), w["f1"]))
            g64_all.append(fused[0]), g8_all.append(fused[1])
            # SpatialEncoder rel_z_decay (A.5)
            Rt = self.extrin[v]
            c = _t(affine_rows(pts, Rt))
            kc = _t(affine_rows(self.kpt3d, Rt))
            dz = 1.0 * (c[:, None, 2] - kc[None, :, 2])                                     # (N,42)
            dxyz = c[:, None] - kc[None]
            wgt = torch.exp(-(dxyz ** 2).sum(-1) / (2.0 * (0.1 ** 2)))                      # (N,42)
            rows = [dz]
            for l in range(3):
                fq = float(np.float32(np.pi * (2 ** l)))
                rows += [torch.sin(dz * fq), torch.cos(dz * fq)]
            pe = (torch.stack(rows, 1) * wgt[:, None]).reshape(N, -1)                       # (N,294)
            pe_all.append(pe)
            h = softplus100(F.linear(torch.cat([pe, fused[0]], 1), *net.mlp[0]))
            h = softplus100(F.linear(h, *net.mlp[1]))
            h = softplus100(F.linear(torch.cat([h, fused[1]], 1), *net.mlp[2]))
            h = F.linear(h, *net.mlp[3])
            lat_views.append(h)
        hv = torch.stack(lat_views)                                                        # (V,N,64)
        wv = _t(pw)[:, :, None]
        mean = (wv * hv).sum(0)
        var = (wv * (hv - mean[None]).pow(2.0)).sum(0)
        latent = torch.cat([mean, var], 1)                                                 # (N,128)
        o = softplus100(F.linear(latent, *net.post[0]))
        o = softplus100(F.linear(o, *net.post[1]))
        o = F.linear(o, *net.post[2])                                                      # (N,2)

        # ---- query_color (A.8)
        lat24 = F.linear(latent, *net.compress)
        viewt = _t(view)
        feats, rdiffs = [], []
        for v in range(V):
            vis = fr["vert_vis"][v]
            qv = _t(geo["qvis"][v].astype(f32))[:, None]
            vn, vt = _t(vis[nn])[:, None], _t(vis[twin])[:, None]
            q = torch.cat([_t(bilinear_np(self.img[v], xy_all[v])), _t(bilinear_np(self.tex[v], xy_all[v]))], 1)   # 11
            tab = fr["Ttex"][v]
            a, b = _t(tab[nn]) * vn, _t(tab[twin]) * vt
            a11, a18, b11, b18 = a[:, :11], a[:, 11:], b[:, :11], b[:, 11:]
            y = torch.cat([q, a11, b11, a18, b18, lat24, qv, vn, vt], 1)                                           # 96
            w = net.tex
            att = torch.sigmoid(F.linear(F.relu(F.linear(y, w["at0"])), w["at1"]))
            y2 = torch.cat([q * att[:, 0:1], a11 * att[:, 1:2], b11 * att[:, 2:3], a18 * att[:, 3:4],
                            b18 * att[:, 4:5], lat24 * att[:, 5:6], qv, vn, vt], 1)
            feats.append(F.linear(F.relu(F.linear(y2, w["f0"])), w["f1"]))                                          # (N,40)
            sp = _t(fr["src_pos"][v])[None]
            sdir = F.normalize(_t(pts) - sp, p=2, dim=-1)
            rd = viewt - sdir
            rn = torch.norm(rd, dim=-1, keepdim=True)
            dot = (sdir * viewt).sum(-1, keepdim=True)
            rdiffs.append(torch.cat([rd / torch.clamp(rn, min=1e-6), dot], -1))
        rgb_feat = torch.stack(feats, 1)                                    # (N,V,40)
        ray_diff = torch.stack(rdiffs, 1)                                   # (N,V,4)
        pmask = _t(out_mask.T.copy())[:, :, None]                           # (N,V,1)
        rgb = self.ibr_head(rgb_feat, ray_diff, pmask)
        out = torch.cat([o, rgb], 1).numpy()
        if taps is not None:
            taps.update(xy=np.stack(xy_all), zn=np.stack(zn_all), in_mask=in_all, fg=fg_all, out_mask=m,
                        pix_weight=pw, pe=torch.stack(pe_all).numpy(), geo64=torch.stack(g64_all).numpy(),
                        geo8=torch.stack(g8_all).numpy(), h3=hv.numpy(), latent=latent.numpy(), lat24=lat24.numpy(),
                        rgb_feat=rgb_feat.numpy(), ray_diff=ray_diff.numpy(), o=o.numpy(), rgb=rgb.numpy())
        return out, valid

    def ibr_head(self, rgb_feats, ray_diffs, mask):
        """IBRRenderingHead.forward (model.py:1600-1636); tensors (N,V,C)."""
        p = self.net.ibr
        elu = F.elu
        V = rgb_feats.shape[1]
        d = elu(F.linear(elu(F.linear(ray_diffs, *p["ray"][0])), *p["ray"][1]))
        src_rgb = rgb_feats[..., :3]
        f = rgb_feats + d
        dot = ray_diffs[..., 3:4]
        e = torch.exp(torch.abs(p["ani_al"]) * (dot - 1))
        wgt = (e - torch.min(e, dim=1, keepdim=True)[0]) * mask
        wgt = wgt / (torch.sum(wgt, dim=1, keepdim=True) + 1e-8)
        mean = torch.sum(f * wgt, dim=1, keepdim=True)
        var = torch.sum(wgt * (f - mean) ** 2, dim=1, keepdim=True)
        x = torch.cat([mean.expand(-1, V, -1), var.expand(-1, V, -1), f], -1)
        x = elu(F.linear(elu(F.linear(x, *p["base"][0])), *p["base"][1]))
        pv = elu(F.linear(elu(F.linear(x * wgt, *p["vis1"][0])), *p["vis1"][1]))
        res, vis = pv[..., :-1], pv[..., -1:]
        x = x + res
        vis = torch.sigmoid(F.linear(elu(F.linear(x * torch.sigmoid(vis) * mask, *p["vis2"][0])), *p["vis2"][1])) * mask
        s = F.linear(elu(F.linear(elu(F.linear(torch.cat([x, vis, ray_diffs], -1), *p["out"][0])), *p["out"][1])), *p["out"][2])
        s = s.masked_fill(mask == 0, -1e4)
        return torch.sum(src_rgb * torch.softmax(s, dim=1), dim=1)

    # ---------------------------------------------------------------- eval_func (A.9, model.py:1140-1160)
    def eval_rgba(self, pts, view, geo, n_samples, noise=None, taps=None):
        out, valid = self.query(pts, view, geo, n_samples, taps)
        m = valid.astype(f32)
        nml = f32(0.1 / 100.0)
        sdf = m * out[:, 0] + (f32(1.0) - m) * nml
        rad = out[:, 1]
        if noise is not None:
            rad = rad + noise
        alpha = m * np.maximum(rad, f32(0))
        return np.concatenate([alpha[:, None], sdf[:, None], out[:, 2:]], 1).astype(f32), valid

    def rgba2out(self, rgba, z, mesh_sdf):
        """rgba (R,S,5), z (R,S), mesh_sdf (R,S).  Sequential products like torch.cumprod on CPU."""
        beta = self.net.sigmoid_beta
        a = _t(rgba[..., 0] + mesh_sdf)
        sigma = torch.sigmoid(-a / beta) / beta
        zt = _t(z)
        dist = torch.cat([zt[:, 1:] - zt[:, :-1], 1e10 * torch.ones_like(zt[:, :1])], -1)
        c = 1.0 - torch.exp(-sigma * dist)
        contrib = c * torch.cumprod(torch.cat([torch.ones_like(c[:, :1]), 1 - c[:, :-1]], -1), -1)
        rgb = _t(rgba[..., 2:])
        color = (rgb * contrib[..., None]).sum(-2)
        alpha = contrib.sum(-1)
        sdf = (_t(rgba[..., 1]) * contrib).sum(-1) / (alpha + 1e-8)
        depth = (zt * contrib).sum(-1) / (alpha + 1e-8)
        return dict(color=color.numpy(), depth=depth.numpy(), alpha=alpha.numpy(), contrib=contrib.numpy(), sdf=sdf.numpy())

    @staticmethod
    def importance_sample(contrib_inner, z_mid, n_fine, u=None):
        """model.py:1425-1462, uniform branch (u = linspace) unless `u` (R,n_fine) is given.  Sequential cumsum."""
        c = (contrib_inner + f32(1e-5)).astype(f32)
        tot = np.zeros(c.shape[0], f32)
        for i in range(c.shape[1]):                      # normaliser summed left to right (defined order)
            tot = tot + c[:, i]
        pdf = c / tot[:, None]
        cdf = np.zeros((c.shape[0], c.shape[1] + 1), f32)
        run = np.zeros(c.shape[0], f32)
        for i in range(c.shape[1]):
            run = run + pdf[:, i]
            cdf[:, i + 1] = run
        if u is None:
            u = np.broadcast_to(torch.linspace(0.0, 1.0, steps=n_fine).numpy()[None], (c.shape[0], n_fine))
        u = np.ascontiguousarray(u, f32)
        idx = _n(torch.searchsorted(_t(cdf), _t(u), right=True))
        lo = np.clip(idx - 1, 0, None)
        hi = np.clip(idx, None, cdf.shape[1] - 1)
        cl, ch = np.take_along_axis(cdf, lo, 1), np.take_along_axis(cdf, hi, 1)
        zl, zh = np.take_along_axis(z_mid, lo, 1), np.take_along_axis(z_mid, hi, 1)
        num = u - cl
        den = ch - cl
        den = np.where(den < f32(1e-5), f32(1.0), den)
        return (zl + (num / den) * (zh - zl)).astype(f32)

    # ---------------------------------------------------------------- batch_render_pifu_nerf (inference, uniform)
    def render(self, level=1, stride_xy=(0, 0), S_c=64, S_f=64, fine=True, pixels=None, taps: Optional[dict] = None):
        cam_tar, inp = self.inp["cam_tar"], self.inp
        if self.frame is None:
            self.frame_setup()
        if pixels is None:
            grids, index = pixel_grid(self.H, self.W, level, stride_xy)
        else:
            grids = np.asarray(pixels, np.int64)
            index = grids[:, 0] + grids[:, 1] * self.W
        inv_K, R, cam_pos = frame_camera(cam_tar)
        rays = make_rays(grids, inv_K, R, cam_pos, cam_tar.get("znear", self.znear), cam_tar.get("zfar", self.zfar),
                         _n(inp["bounds"])[0])
        t_tab = torch.linspace(0.0, 1.0, steps=S_c).numpy().astype(f32)
        z = sample_z(rays["znear"], rays["zfar"], t_tab)
        pts = sample_points(cam_pos, rays["dirs"], z)
        Rn = grids.shape[0]
        view = np.repeat(rays["dirs"], S_c, 0)
        geo = self.geometry(pts)
        tq = {} if taps is not None else None
        rgba, valid = self.eval_rgba(pts, view, geo, S_c, taps=tq)
        comp = self.rgba2out(rgba.reshape(Rn, S_c, 5), z, geo["sdf"].reshape(Rn, S_c))
        out = dict(index=index, tex_fg=comp["color"], depth=comp["depth"], alpha=comp["alpha"])
        if taps is not None:
            taps.update(rays=rays, z=z, pts=pts, geo=geo, query=tq, rgba=rgba, valid=valid, contrib=comp["contrib"],
                        cam_pos=cam_pos, inv_K=inv_K, R=R, t_tab=t_tab)
        if fine:
            z_mid = (f32(0.5) * (z[:, 1:] + z[:, :-1])).astype(f32)
            z_f = self.importance_sample(comp["contrib"][:, 1:-1], z_mid, S_f)
            z2 = np.sort(np.concatenate([z, z_f], 1), 1).astype(f32)
            pts2 = sample_points(cam_pos, rays["dirs"], z2)
            view2 = np.repeat(rays["dirs"], z2.shape[1], 0)
            geo2 = self.geometry(pts2)
            tq2 = {} if taps is not None else None
            rgba2, valid2 = self.eval_rgba(pts2, view2, geo2, S_f, taps=tq2)
            comp2 = self.rgba2out(rgba2.reshape(Rn, -1, 5), z2, geo2["sdf"].reshape(Rn, -1))
            out.update(tex_fg_fine=comp2["color"], depth_fine=comp2["depth"], alpha_fine=comp2["alpha"], sdf=comp2["sdf"])
            if taps is not None:
                taps.update(z_fine_only=z_f, z_fine=z2, pts_fine=pts2, geo_fine=geo2, query_fine=tq2, rgba_fine=rgba2,
                            valid_fine=valid2, contrib_fine=comp2["contrib"])
        out["vert_vis"] = self.frame["vert_vis"]
        return out
