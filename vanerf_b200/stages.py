"""Stage-level callables of the reference with their names, argument meaning and return layouts (SURVEY.md §8(b) "kept
surface"), computed by the stage primitives of libvanerf_b200.so (csrc/stages.cuh: vanerf_feat_sample, vanerf_knn1,
vanerf_dense, vanerf_rel_z_decay):

    feat_sample(feat, uv)                                  src/utils.py:136-151
    KNN_vis(query, vert, vert_feat, vert_vis, k)           src/networks.py:27-33
    SpatialEncoder(...).forward(**sp_data) / .position_embedding / .get_dim      src/spatial.py:20-117
    GeoVisFusion(...).forward(vert_xy, fg, feat_sampled, vert, v, vert_vis, query_vis, closest_face, query_sdf)   src/networks.py:75-106
    TexVisFusion(...).forward(vert_xy, ft1, ft_xy, vert, v, vert_vis, query_vis, img_xy, img_fmap, latent_fused)  src/networks.py:268-293
    MLPUNetFusion(...).forward(x, f, a, w)                 src/utils.py:633-649
    IBRRenderingHead(...).forward(rgb_feats, ray_diffs, proj_mask)               src/model.py:1600-1636

They exist for callers (and tests) that use one stage of the reference on its own.  The render path
(`vanerf_b200.model.VANeRF` -> `Renderer` -> torch.ops.vanerf_b200.*) never calls them: the fused kernels compute the
same stages without materialising their outputs.  Dense layers, gathers, the nearest-vertex search and the positional
encoding run in the library; torch does the reshapes, concatenations and the per-element gate / pooling arithmetic
between them.  Weights are taken from a reference `state_dict` (same keys).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib as L
from . import weights as W
from .renderer import Renderer

NUM_V = 1558 // 2          # src/networks.py:25
ACT = {"none": 0, "relu": 1, "softplus": 2, "sigmoid": 3, "elu": 4}


class _Backend:
    """Binds the stage primitives on one device (owns a library context for stream / error plumbing)."""

    def __init__(self, device="cuda:0", lib: Optional[L.Lib] = None, renderer: Optional[Renderer] = None):
        self.r = renderer or Renderer(device, lib)
        self.device = self.r.device

    def _f(self, t):
        return t.detach().to(self.device, torch.float32).contiguous()

    def feat_sample(self, feat, uv):
        feat, uv = self._f(feat), self._f(uv)
        B, Cc, H, Wd = feat.shape
        N = uv.shape[1]
        out = self.r.empty((B, N, Cc))
        st = self.r.lib.dll.vanerf_feat_sample(self.r.ctx, self.r._ptr(feat), B, Cc, H, Wd, self.r._ptr(uv), N, self.r._ptr(out), self.r.stream)
        self.r.lib.check(self.r.ctx, st, "vanerf_feat_sample")
        return out

    def knn1(self, query, vert):
        q, v = self._f(query).reshape(-1, 3), self._f(vert).reshape(-1, 3)
        idx = self.r.empty((q.shape[0],), torch.int32)
        st = self.r.lib.dll.vanerf_knn1(self.r.ctx, self.r._ptr(q), q.shape[0], self.r._ptr(v), v.shape[0], self.r._ptr(idx), self.r.stream)
        self.r.lib.check(self.r.ctx, st, "vanerf_knn1")
        return idx.long()

    def dense(self, x, w, b=None, act="none"):
        """x (..., K) -> act(x @ w^T + b) (..., N)."""
        x = self._f(x)
        lead, K = x.shape[:-1], x.shape[-1]
        x2 = x.reshape(-1, K)
        w = self._f(w)
        bb = self._f(b) if b is not None else None
        y = self.r.empty((x2.shape[0], w.shape[0]))
        st = self.r.lib.dll.vanerf_dense(self.r.ctx, self.r._ptr(x2), x2.shape[0], K, self.r._ptr(w), self.r._ptr(bb), w.shape[0], ACT[act],
                                         self.r._ptr(y), self.r.stream)
        self.r.lib.check(self.r.ctx, st, "vanerf_dense")
        return y.reshape(*lead, w.shape[0])

    def rel_z_decay(self, cxyz, kxyz, levels, scale, sigma):
        c, k = self._f(cxyz), self._f(kxyz)
        BV, N, Kp = c.shape[0], c.shape[1], k.shape[1]
        out = self.r.empty((BV, N, (1 + 2 * levels) * Kp))
        st = self.r.lib.dll.vanerf_rel_z_decay(self.r.ctx, self.r._ptr(c), self.r._ptr(k), BV, N, Kp, levels, float(scale), float(sigma),
                                               self.r._ptr(out), self.r.stream)
        self.r.lib.check(self.r.ctx, st, "vanerf_rel_z_decay")
        return out


_default: Dict[str, _Backend] = {}


def backend(device="cuda:0", lib: Optional[L.Lib] = None) -> _Backend:
    key = f"{device}|{lib.path if lib else ''}"
    if key not in _default:
        _default[key] = _Backend(device, lib)
    return _default[key]


# ---------------------------------------------------------------------------------------------------- functions
def feat_sample(feat, uv, mode="bilinear", padding_mode="border", align_corners=True, be: Optional[_Backend] = None):
    """src/utils.py:136-151: feat (B,C,H,W), uv (B,N,2) in [-1,1] -> (B,N,C)."""
    assert (mode, padding_mode, align_corners) == ("bilinear", "border", True), "the path uses bilinear / border / align_corners=True only"
    return (be or backend()).feat_sample(feat, uv)


def KNN_vis(query, vert, vert_feat, vert_vis, k=1, be: Optional[_Backend] = None):
    """src/networks.py:27-33: nearest vertex (K=1, batch element 0's indices for every row, as the reference does) -> its
    feature row x visibility, the twin vertex's ((id + 779) mod 1558) row x visibility, and the two visibilities."""
    assert k == 1
    be = be or backend()
    idx = be.knn1(query[0], vert[0])                                       # mink_idxs[0, :, 0]
    vf, vv = vert_feat.to(be.device).float(), vert_vis.to(be.device).float()
    nv = vf.shape[1] // 2
    twin = (idx + nv) % vf.shape[1]
    a = vf[:, idx] * vv[:, idx]
    b = vf[:, twin] * vv[:, twin]
    return a, b, vv[:, idx], vv[:, twin]


# ---------------------------------------------------------------------------------------------------- modules
class _Stage:
    def __init__(self, state_dict=None, device="cuda:0", lib: Optional[L.Lib] = None, be: Optional[_Backend] = None):
        self.be = be or backend(device, lib)
        self.w: Dict[str, np.ndarray] = {}
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def load_state_dict(self, state_dict):
        self.w = W.fold(W.strip_prefix(state_dict))
        return self

    def _lin(self, x, tag, act="none"):
        return self.be.dense(x, torch.from_numpy(np.ascontiguousarray(self.w[tag + ".w"])),
                             torch.from_numpy(np.ascontiguousarray(self.w[tag + ".b"])) if (tag + ".b") in self.w else None, act)

    def __call__(self, *a, **k):
        return self.forward(*a, **k)


class SpatialEncoder(_Stage):
    """src/spatial.py.  sp_type "rel_z_decay" (configs/vanerf.json) runs in the library; the parameter-free helper
    `position_embedding` / `get_dim` keep the reference's static semantics."""

    def __init__(self, sp_level=3, sp_type="rel_z_decay", scale=1.0, n_kpt=42, sigma=0.1, device="cuda:0", lib=None, be=None, **kwargs):
        super().__init__(None, device, lib, be)
        self.sp_level, self.sp_type, self.scale, self.n_kpt, self.sigma = sp_level, sp_type, scale, n_kpt, kwargs.get("sigma", sigma)

    @staticmethod
    def pe_vector(nlevels, device, scale=1.0):
        v, val = [], 1
        for _ in range(nlevels):
            v.append(scale * np.pi * val)
            val *= 2
        return torch.from_numpy(np.asarray(v, dtype=np.float32)).to(device)

    @staticmethod
    def position_embedding(x, nlevels, scale=1.0):
        """(B,N,C) -> (B,N,C (1 + 2 nlevels)): [x | sin(x f_0) | cos(x f_0) | ...] (src/spatial.py:20-35); plain torch, like the reference."""
        if nlevels <= 0:
            return x
        vec = SpatialEncoder.pe_vector(nlevels, x.device, scale)
        B, N, _ = x.shape
        y = x[:, :, None, :] * vec[None, None, :, None]
        z = torch.cat((torch.sin(y), torch.cos(y)), axis=-1).view(B, N, -1)
        return torch.cat([x, z], -1)

    def get_dim(self):
        if self.sp_type in ["z", "rel_z", "rel_z_decay"]:
            return (1 + 2 * self.sp_level) * self.n_kpt if "rel" in self.sp_type else 1 + 2 * self.sp_level
        if "xyz" in self.sp_type:
            return (1 + 2 * self.sp_level) * 3 * (self.n_kpt if "rel" in self.sp_type else 1)
        return 0

    def forward(self, **sp_data):
        if self.sp_type != "rel_z_decay":
            raise NotImplementedError("only sp_type='rel_z_decay' (configs/vanerf.json) is on the path")
        dev = self.be.device
        v, Rt, kpt = sp_data["v"].to(dev).float(), sp_data["extrin"].to(dev).float(), sp_data["kpt3d"].to(dev).float()
        V = sp_data["n_view"]
        assert kpt.shape[1] == self.n_kpt
        cxyz = v @ Rt[:, :3, :3].transpose(1, 2) + Rt[:, :3, 3][:, None]
        k3 = kpt[:, None].expand(-1, V, -1, -1).reshape(-1, *kpt.shape[1:])
        kxyz = k3 @ Rt[:, :3, :3].transpose(1, 2) + Rt[:, :3, 3][:, None]
        return self.be.rel_z_decay(cxyz, kxyz, self.sp_level, self.scale, self.sigma)


class GeoVisFusion(_Stage):
    """src/networks.py:75-106 (two scales, 64 and 8 channels)."""

    def forward(self, vert_xy, fg, feat_sampled, vert, v, vert_vis, query_vis, closest_face, query_sdf):
        be = self.be
        dev = be.device
        B = vert_xy.shape[0]
        qs, qv = query_sdf.to(dev).float(), query_vis.to(dev).float()
        out = []
        for s, (at, f) in enumerate([("geo_at", "geo_f"), ("geo8_at", "geo8_f")]):
            vf = be.feat_sample(fg[s], vert_xy)
            a, b, va, vb = KNN_vis(v, vert, vf, vert_vis, 1, be)
            px = feat_sampled[s].to(dev).float().reshape(B, -1, vf.shape[-1])
            x = torch.cat([px, a, b, qs, qv, va, vb], 2)
            g = self._lin(self._lin(x, at + "0", "relu"), at + "1", "sigmoid")
            y = torch.cat([px * g[:, :, 0:1], a * g[:, :, 1:2], b * g[:, :, 2:3], qs, qv, va, vb], 2)
            y = self._lin(self._lin(y, f + "0", "relu"), f + "1")
            out.append(y.view(B, 1, *y.shape[-2:]))
        return out


class TexVisFusion(_Stage):
    """src/networks.py:268-293.  The per-frame global feature (fconv3 / fconv4 / fconv_gt, :273-279) is produced by the
    renderer's per-frame setup; pass it as `gf_vert_feat` (BV,1558,18), or pass a `Renderer` with loaded weights to compute it."""

    def __init__(self, state_dict=None, device="cuda:0", lib=None, be=None, renderer: Optional[Renderer] = None):
        super().__init__(state_dict, device, lib, be)
        self.renderer = renderer

    def forward(self, vert_xy, ft1, ft_xy, vert, v, vert_vis, query_vis, img_xy, img_fmap, latent_fused, gf_vert_feat=None):
        be = self.be
        dev = be.device
        vf = torch.cat([be.feat_sample(img_fmap, vert_xy), be.feat_sample(ft1, vert_xy)], 2)
        if gf_vert_feat is None:
            assert self.renderer is not None, "gf_vert_feat or a Renderer with weights is needed for the global vertex feature"
            gf_vert_feat = self.renderer.global_vertex_feature(img_fmap.to(dev).float(), ft1.to(dev).float())
        vf = torch.cat([vf, gf_vert_feat.to(dev).float()], 2)
        a, b, va, vb = KNN_vis(v, vert, vf, vert_vis, 1, be)
        a_gf, b_gf, a, b = a[:, :, 11:], b[:, :, 11:], a[:, :, :11], b[:, :, :11]
        q = torch.cat([img_xy.to(dev).float(), ft_xy.to(dev).float()], 2)
        lat, qv = latent_fused.to(dev).float(), query_vis.to(dev).float()
        y = torch.cat([q, a, b, a_gf, b_gf, lat, qv, va, vb], 2)
        g = self._lin(self._lin(y, "tex_at0", "relu"), "tex_at1", "sigmoid")
        y = torch.cat([q * g[:, :, 0:1], a * g[:, :, 1:2], b * g[:, :, 2:3], a_gf * g[:, :, 3:4], b_gf * g[:, :, 4:5], lat * g[:, :, 5:6],
                       qv, va, vb], 2)
        return self._lin(self._lin(y, "tex_f0", "relu"), "tex_f1")


class MLPUNetFusion(_Stage):
    """src/utils.py:633-649 with configs/vanerf.json's shapes: layers1 358-128-128-(136)120-64 (skip at layer 0: f[0], layer 2:
    f[1]), weighted mean + variance pooling over views, layers2 128-64-64-2."""

    def forward(self, x, f: List[torch.Tensor], a, w=None, x_add=None, nonlin=None):
        dev = self.be.device
        x, a = x.to(dev).float(), a.to(dev).float()
        f = [t.to(dev).float() for t in f]
        h = self._lin(torch.cat([x, f[0]], -1), "mlp0", "softplus")
        h = self._lin(h, "mlp1", "softplus")
        h = self._lin(torch.cat([h, f[1]], -1), "mlp2", "softplus")
        x_view = self._lin(h, "mlp3")
        a_sum = a.sum(1)
        w = a / (a_sum[:, None] + 1e-6) if w is None else w.to(dev).float()
        mean = (w * x_view).sum(1)
        var = (w * (x_view - mean[:, None]).pow(2.0)).sum(1)
        x_pool = torch.cat([mean, var], -1)
        valid = a_sum > 0.0
        if x_add is not None:
            x_pool = torch.cat([x_pool, x_add.to(dev).float()], -1)
        out = self._lin(self._lin(self._lin(x_pool, "post0", "softplus"), "post1", "softplus"), "post2")
        return out, valid, x_view, x_pool


class IBRRenderingHead(_Stage):
    """src/model.py:1600-1636."""

    def forward(self, rgb_feats, ray_diffs, proj_mask):
        dev = self.be.device
        rgb_feats, ray_diffs, proj_mask = rgb_feats.to(dev).float(), ray_diffs.to(dev).float(), proj_mask.to(dev).float()
        V = rgb_feats.shape[2]
        dir_feat = self._lin(self._lin(ray_diffs, "ray0", "elu"), "ray1", "elu")
        src_rgb = rgb_feats[..., :3]
        rgb_feats = torch.cat((rgb_feats[..., :dir_feat.shape[-1]] + dir_feat, rgb_feats[..., dir_feat.shape[-1]:]), dim=-1)
        dot_prod = ray_diffs[..., 3:4]
        e = torch.exp(abs(float(self.w["ani_al"][0])) * (dot_prod - 1))
        weight = (e - torch.min(e, dim=2, keepdim=True)[0]) * proj_mask
        weight = weight / (torch.sum(weight, dim=2, keepdim=True) + 1e-8)
        mean = torch.sum(rgb_feats * weight, dim=2, keepdim=True)
        var = torch.sum(weight * (rgb_feats - mean) ** 2, dim=2, keepdim=True)
        fused = torch.cat([mean, var], dim=-1)
        x = self._lin(self._lin(torch.cat([fused.expand(-1, -1, V, -1), rgb_feats], dim=-1), "base0", "elu"), "base1", "elu")
        pv = self._lin(self._lin(x * weight, "vis10", "elu"), "vis11", "elu")
        res, vis = pv[..., :-1], pv[..., -1:]
        x = x + res
        vis = self._lin(self._lin(x * torch.sigmoid(vis) * proj_mask, "vis20", "elu"), "vis21", "sigmoid") * proj_mask
        s = self._lin(self._lin(self._lin(torch.cat([x, vis, ray_diffs], dim=-1), "outl0", "elu"), "outl1", "elu"), "outl2")
        s = s.masked_fill(proj_mask == 0, -1e4)
        return torch.sum(src_rgb * torch.softmax(s, dim=2), dim=2)
