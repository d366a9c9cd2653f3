"""Training branch of the render path (BASELINE.json configs[4]: forward + backward through the path, gradient all-reduce).

Reference: the `net.training` paths of `VANeRF.batch_render_pifu_nerf` / `VANeRF.query` (src/model.py:748-957, :1103-1422) —
random 64x64 patch (:1172-1189), stratified depth jitter (:1226-1230), view dropout (:804-810), density noise (:1155-1156),
random importance samples (:1439-1442) — their torch autograd, and the DDP gradient all-reduce of train.py:58,65.

What runs where
  * our kernels (libvanerf_b200.so): ray generation + box clip + depths (`vanerf_sample_rays_t`, bit-exact), importance
    sampling with per-ray u (`vanerf_importance`, bit-exact given contrib), mesh queries (`vanerf_geom_query`), projection /
    masks / boundary weights / camera-space positions / ray differences (`vanerf_project_samples`), the positional
    encoding (`vanerf_rel_z_decay`), the bilinear gathers forward and backward (`vanerf_feat_sample`, `_bwd`: scatter-add
    into the feature maps) and alpha compositing forward and backward (`vanerf_composite_beta`, `_bwd`);
  * library GEMMs under torch autograd: the dense layers of the UNFUSED training graph (`torch.nn.functional.linear`, i.e.
    cuBLAS), the per-frame TexVisFusion convolution stacks (cuDNN) and the element-wise glue between them.  The fused
    tcgen05 / FFMA kernels of the inference path keep no activations and are not differentiated.
Parameters live under the reference's `state_dict` names, so an optimiser / checkpoint of the reference applies.

Randomness: the reference draws from torch's / numpy's global generators at several places.  `TrainRandom`
draws the same tensors in the same order from explicit generators (on the CPU, then moved to the device), so a parity
harness can seed both sides identically; without a seed it draws fresh numbers like the reference.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L
from .renderer import Renderer

NUM_V = 779


# ---------------------------------------------------------------------------------------------------- autograd functions
class FeatSampleFn(torch.autograd.Function):
    """feat_sample (src/utils.py:136-151) with a hand-written backward w.r.t. the map (scatter-add kernel)."""

    @staticmethod
    def forward(ctx, r: Renderer, feat: torch.Tensor, uv: torch.Tensor):
        feat = feat.contiguous()
        uv = uv.detach().contiguous()
        B, Cc, H, Wd = feat.shape
        N = uv.shape[1]
        out = r.empty((B, N, Cc))
        st = r.lib.dll.vanerf_feat_sample(r.ctx, r._ptr(feat.detach()), B, Cc, H, Wd, r._ptr(uv), N, r._ptr(out), r.stream)
        r.lib.check(r.ctx, st, "vanerf_feat_sample")
        ctx.r, ctx.shape = r, (B, Cc, H, Wd)
        ctx.save_for_backward(uv)
        return out

    @staticmethod
    def backward(ctx, g):
        (uv,) = ctx.saved_tensors
        r = ctx.r
        B, Cc, H, Wd = ctx.shape
        d_feat = torch.zeros((B, Cc, H, Wd), dtype=torch.float32, device=g.device)
        g = g.contiguous()
        st = r.lib.dll.vanerf_feat_sample_bwd(r.ctx, r._ptr(g), B, Cc, H, Wd, r._ptr(uv), uv.shape[1], r._ptr(d_feat), r.stream)
        r.lib.check(r.ctx, st, "vanerf_feat_sample_bwd")
        return None, d_feat, None


class CompositeFn(torch.autograd.Function):
    """VANeRF.rgba2out (src/model.py:1465-1494): forward and backward in the compositing kernels; beta is a parameter."""

    @staticmethod
    def forward(ctx, r: Renderer, rgba: torch.Tensor, z: torch.Tensor, mesh_sdf: torch.Tensor, beta: torch.Tensor):
        rgba, z, mesh_sdf = rgba.contiguous(), z.contiguous(), mesh_sdf.contiguous()
        R, S = z.shape
        b = float(beta.detach().reshape(-1)[0].item())
        color, depth, alpha, sdf, contrib = r.empty((R, 3)), r.empty((R,)), r.empty((R,)), r.empty((R,)), r.empty((R, S))
        st = r.lib.dll.vanerf_composite_beta(r.ctx, r._ptr(rgba.detach()), r._ptr(z), r._ptr(mesh_sdf), R, S, b, r._ptr(color), r._ptr(depth),
                                             r._ptr(alpha), r._ptr(sdf), r._ptr(contrib), r.stream)
        r.lib.check(r.ctx, st, "vanerf_composite_beta")
        ctx.r, ctx.b = r, b
        ctx.save_for_backward(rgba.detach(), z, mesh_sdf)
        ctx.mark_non_differentiable(contrib)
        return color, depth, alpha, sdf, contrib

    @staticmethod
    def backward(ctx, g_color, g_depth, g_alpha, g_sdf, _g_contrib):
        rgba, z, mesh_sdf = ctx.saved_tensors
        r = ctx.r
        R, S = z.shape
        d_rgba = torch.zeros((R, S, 5), dtype=torch.float32, device=z.device)
        d_beta = torch.zeros((1,), dtype=torch.float32, device=z.device)
        # contiguous copies of the incoming gradients must stay referenced until the call has been issued: a pointer taken from a
        # temporary dangles as soon as the temporary is released (its block is handed to the next allocation)
        keep = [t.contiguous() if t is not None else None for t in (g_color, g_alpha, g_depth, g_sdf)]
        p = lambda t: r._ptr(t) if t is not None else None
        st = r.lib.dll.vanerf_composite_bwd(r.ctx, r._ptr(rgba), r._ptr(z), r._ptr(mesh_sdf), R, S, ctx.b, p(keep[0]), p(keep[1]), p(keep[2]),
                                            p(keep[3]), r._ptr(d_rgba), r._ptr(d_beta), r.stream)
        r.lib.check(r.ctx, st, "vanerf_composite_bwd")
        return None, d_rgba, None, None, d_beta


TF32 = False        # module switch set by `matmul_precision`: the training graph's GEMMs and convolutions on TF32 tensor cores


def fp32_exact():
    """Context for forward AND backward of the training graph: cuDNN / cuBLAS in true fp32 (no TF32), deterministic algorithms.
    The reference's own GPU run takes TF32 for its convolutions; the CPU oracle is the arbiter here.  The
    backward of a convolution reads these flags when it RUNS, so `loss.backward()` belongs inside the context too.
    Inside `matmul_precision("tf32")` the convolutions follow the GEMMs onto the tensor cores."""
    return torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=not TF32, allow_tf32=TF32)


class matmul_precision:
    """`with matmul_precision("tf32"):` runs the dense layers of the unfused training graph (cuBLAS GEMMs, cuDNN convolutions) on TF32
    tensor cores (10-bit mantissa operands, fp32 accumulation) for the steps inside it; "fp32" (the default everywhere, and what the
    gradient-parity tests run) keeps them exact.  An opt-in like the bf16-MLP path of inference: faster, not bit-comparable."""

    def __init__(self, mode: str = "fp32"):
        if mode not in ("fp32", "tf32"):
            raise ValueError("matmul_precision: 'fp32' or 'tf32'")
        self.on = mode == "tf32"

    def __enter__(self):
        global TF32
        self.prev = (TF32, torch.backends.cuda.matmul.allow_tf32)
        TF32 = self.on
        torch.backends.cuda.matmul.allow_tf32 = self.on
        return self

    def __exit__(self, *exc):
        global TF32
        TF32, torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


# ---------------------------------------------------------------------------------------------------- randomness
class TrainRandom:
    """The random tensors of one training forward, drawn in the reference's order and shapes:
    patch centre (numpy), coarse jitter rand_like(z) (B,R,S_c), then per query: view-dropout rand_like (B,V-1,1,1) and
    rand_like (B,V,1,1), density noise randn_like (B,N,1); between the passes the importance samples rand (B,R,S_f)."""

    def __init__(self, seed: Optional[int] = None, np_seed: Optional[int] = None):
        self.gen = torch.Generator(device="cpu")
        if seed is None:
            self.gen.seed()
        else:
            self.gen.manual_seed(seed)
        self.np = np.random.RandomState(np_seed) if np_seed is not None else np.random

    def patch_center_index(self, n: int) -> int:
        return int(self.np.randint(0, n, 1)[0])

    def rand(self, *shape) -> torch.Tensor:
        return torch.rand(*shape, generator=self.gen)

    def randn(self, *shape) -> torch.Tensor:
        return torch.randn(*shape, generator=self.gen)

    def view_dropout(self, V: int) -> torch.Tensor:
        """src/model.py:804-810 -> (V,1) float {0,1}: `dropout = zeros_like(out_mask[:, :, :1])` has shape (B,V,1,1), i.e. whole
        views are dropped for the entire ray batch.  Slot 0 is kept, the others with p = 0.5, then the slots are permuted by the
        argsort of a second uniform draw."""
        d = torch.zeros(1, V, 1, 1)
        d[:, :1] = 1.0
        d[:, 1:] = (self.rand(1, V - 1, 1, 1) > 0.5).float()
        perm = self.rand(1, V, 1, 1).argsort(dim=1)
        return torch.gather(d, 1, perm)[0, :, :, 0]


# ---------------------------------------------------------------------------------------------------- the trainable path
def _softplus100(x):
    return F.softplus(x, beta=100.0, threshold=20.0)


class TrainableRenderPath(torch.nn.Module):
    """Render-path parameters (reference `state_dict` keys) + differentiable forward.  Batch size 1, V source views."""

    def __init__(self, state_dict, device="cuda:0", lib: Optional[L.Lib] = None, rand_noise_std: float = 0.01):
        super().__init__()
        from . import weights as W
        sd = W.strip_prefix(state_dict)
        self.renderer = Renderer(device, lib)
        self.renderer.load_state_dict(sd)                   # inference kernels (sampling / geometry use the frame state only)
        self.dev = self.renderer.device
        self.rand_noise_std = rand_noise_std
        self._names: Dict[str, str] = {}
        for k, v in sd.items():
            if k.startswith(("geo_encoder", "tex_encoder", "vgg_loss", "sp_encoder")):
                continue
            t = torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).detach().float().to(self.dev).clone()
            pname = k.replace(".", "__")
            self._names[k] = pname
            self.register_parameter(pname, torch.nn.Parameter(t))

    # ---- parameter access under the reference names
    def P(self, key: str) -> torch.Tensor:
        return getattr(self, self._names[key])

    def state_dict_ref(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, p).detach().clone() for k, p in self._names.items()}

    def named_grads(self) -> Dict[str, Optional[torch.Tensor]]:
        return {k: getattr(self, p).grad for k, p in self._names.items()}

    def _conv1(self, x, key):                               # Conv1d(k=1, bias=False) on channel-last rows
        return F.linear(x, self.P(key + ".weight")[:, :, 0])

    def _wn(self, x, prefix):                               # weight-normed Linear (src/utils.py:670-685)
        v, g = self.P(prefix + "weight_v"), self.P(prefix + "weight_g")
        return F.linear(x, v * (g / v.norm(dim=1, keepdim=True)), self.P(prefix + "bias"))

    def _lin(self, x, prefix):
        return F.linear(x, self.P(prefix + ".weight"), self.P(prefix + ".bias"))

    # ---- per frame
    def set_frame(self, frame: Dict):
        """frame: reference-layout dictionaries of one time step (img, cam_in, targets, sp_data, feat_geo, feat_tex,
        src_foreground_mask).  Feature maps may require grad (encoders upstream)."""
        r = self.renderer
        dev = self.dev
        f32 = lambda t: t.to(dev, torch.float32)
        self.img, self.g0, self.g1, self.tex = f32(frame["img"]), f32(frame["feat_geo"][0]), f32(frame["feat_geo"][1]), f32(frame["feat_tex"])
        with torch.no_grad():
            self.vert_vis = r.set_frame(self.img.detach(), frame["cam_in"], frame["targets"], frame["sp_data"],
                                        [self.g0.detach(), self.g1.detach()], self.tex.detach(), frame["src_foreground_mask"])   # (V,Nv)
        V = self.img.shape[0]
        cam = frame["cam_in"]
        KRT = cam["KRT"].to(dev).float()
        vert = frame["targets"]["vert_world"].to(dev).float()               # (1,Nv,3)
        vimg = vert @ KRT[:, :3, :3].transpose(1, 2) + KRT[:, :3, 3][:, None]
        vxy = vimg[..., :2] / (vimg[..., 2:3] + 1e-8)
        self.vert_xy = torch.stack([2.0 * (vxy[..., 0] / (float(cam["width"]) - 1.0)) - 1.0,
                                    2.0 * (vxy[..., 1] / (float(cam["height"]) - 1.0)) - 1.0], -1).detach()   # (V,Nv,2)
        extr = frame["sp_data"]["extrin"].to(dev).float()
        kpt = frame["sp_data"]["kpt3d"].to(dev).float()                      # (1,42,3)
        self.kpt_cam = (kpt.expand(V, -1, -1) @ extr[:, :3, :3].transpose(1, 2) + extr[:, :3, 3][:, None]).contiguous()
        self.V = V
        self._tables = None

    def _vertex_tables(self):
        """Differentiable per-frame vertex tables (src/networks.py:83,96,270-279): map taps at the projected vertices and the
        TexVisFusion global feature (conv stacks in torch: their 4 M parameters get gradients through torch autograd)."""
        if self._tables is None:
            r = self.renderer
            T64 = FeatSampleFn.apply(r, self.g0, self.vert_xy)
            T8 = FeatSampleFn.apply(r, self.g1, self.vert_xy)
            sd = {k: self.P(k) for k in self._names if k.startswith(("tex_vis_fusion.fconv3", "tex_vis_fusion.fconv4", "tex_vis_fusion.fconv_gt"))}
            with fp32_exact():
                gf = Renderer._global_vertex_feature(self.img, self.tex, sd)
            Tt = torch.cat([FeatSampleFn.apply(r, self.img, self.vert_xy), FeatSampleFn.apply(r, self.tex, self.vert_xy), gf], 2)
            self._tables = (T64, T8, Tt)
        return self._tables

    # ---- VANeRF.query + eval_func on one set of depths, differentiable
    def shade(self, tar: L.VTarget, rays: torch.Tensor, z: torch.Tensor, drop: Optional[torch.Tensor], noise: Optional[torch.Tensor]):
        r = self.renderer
        V = self.V
        R, S = z.shape
        N = R * S
        geo = r.geom_query(tar, rays, z, want_pts=False)
        xy, cam, rd = r.empty((V, N, 2)), r.empty((V, N, 3)), r.empty((V, N, 4))
        mask, pw_raw = r.empty((N,), torch.uint8), r.empty((V, N))
        st = r.lib.dll.vanerf_project_samples(r.ctx, C.byref(tar), r._ptr(rays), r._ptr(z), R, S, r._ptr(xy), r._ptr(mask), r._ptr(pw_raw),
                                              r._ptr(cam), r._ptr(rd), r.stream)
        r.lib.check(r.ctx, st, "vanerf_project_samples")
        out_mask = mask.float()[None].expand(V, -1)
        if drop is not None:
            out_mask = out_mask * drop.to(self.dev)
        pw = pw_raw * out_mask
        pw = (pw / (pw.sum(0, keepdim=True) + 1e-6))[..., None]              # (V,N,1)
        valid = out_mask.sum(0) > 0
        T64, T8, Tt = self._vertex_tables()
        nn = geo["nn"].long()
        twin = (nn + NUM_V) % (2 * NUM_V)
        vis = self.vert_vis
        vn, vt = vis[:, nn][..., None], vis[:, twin][..., None]              # (V,N,1)
        rows = lambda T, idx: torch.index_select(T, 1, idx)                  # T[:, idx]; its backward is one index_add_ (advanced indexing sorts)
        sdf = geo["sdf"][None, :, None].expand(V, -1, -1)
        qv = geo["qvis"].float()[..., None]
        # ---- GeoVisFusion (src/networks.py:75-106)
        fused = []
        for g, T, at, ff in ((self.g0, T64, "geo_vis_fusion.fconv_at", "geo_vis_fusion.fconv_ated"),
                             (self.g1, T8, "geo_vis_fusion.fconv_at1", "geo_vis_fusion.fconv_ated1")):
            px = FeatSampleFn.apply(r, g, xy)
            a, b = rows(T, nn) * vn, rows(T, twin) * vt
            x = torch.cat([px, a, b, sdf, qv, vn, vt], 2)
            att = torch.sigmoid(self._conv1(F.relu(self._conv1(x, at + ".0")), at + ".2"))
            y = torch.cat([px * att[..., 0:1], a * att[..., 1:2], b * att[..., 2:3], sdf, qv, vn, vt], 2)
            fused.append(self._conv1(F.relu(self._conv1(y, ff + ".0")), ff + ".2"))
        # ---- SpatialEncoder rel_z_decay (no parameters, no gradient)
        pe = r.empty((V, N, 294))
        st = r.lib.dll.vanerf_rel_z_decay(r.ctx, r._ptr(cam), r._ptr(self.kpt_cam), V, N, 42, 3, 1.0, 0.1, r._ptr(pe), r.stream)
        r.lib.check(r.ctx, st, "vanerf_rel_z_decay")
        # ---- MLPUNetFusion (src/utils.py:633-649)
        p1 = "mlp_geo.layers1.layers."
        h = _softplus100(self._wn(torch.cat([pe, fused[0]], 2), p1 + "0.linear."))
        h = _softplus100(self._wn(h, p1 + "1.linear."))
        h = _softplus100(self._wn(torch.cat([h, fused[1]], 2), p1 + "2.linear."))
        h3 = self._lin(h, p1 + "3.linear")
        mean = (pw * h3).sum(0)
        var = (pw * (h3 - mean[None]).pow(2.0)).sum(0)
        latent = torch.cat([mean, var], 1)
        p2 = "mlp_geo.layers2.layers."
        o = self._lin(_softplus100(self._wn(_softplus100(self._wn(latent, p2 + "0.linear.")), p2 + "1.linear.")), p2 + "2.linear")
        # ---- query_color: TexVisFusion + IBRRenderingHead (src/model.py:904-951, networks.py:268-293, model.py:1600-1636)
        lat24 = self._lin(latent, "ibr_compress_gfeat")[None].expand(V, -1, -1)
        q = torch.cat([FeatSampleFn.apply(r, self.img, xy), FeatSampleFn.apply(r, self.tex, xy)], 2)
        a, b = rows(Tt, nn) * vn, rows(Tt, twin) * vt
        a11, a18, b11, b18 = a[..., :11], a[..., 11:], b[..., :11], b[..., 11:]
        y = torch.cat([q, a11, b11, a18, b18, lat24, qv, vn, vt], 2)
        att = torch.sigmoid(self._conv1(F.relu(self._conv1(y, "tex_vis_fusion.fconv_at.0")), "tex_vis_fusion.fconv_at.2"))
        y2 = torch.cat([q * att[..., 0:1], a11 * att[..., 1:2], b11 * att[..., 2:3], a18 * att[..., 3:4], b18 * att[..., 4:5],
                        lat24 * att[..., 5:6], qv, vn, vt], 2)
        rgb_feat = self._conv1(F.relu(self._conv1(y2, "tex_vis_fusion.fconv.0")), "tex_vis_fusion.fconv.2").permute(1, 0, 2)   # (N,V,40)
        rgb = self._ibr_head(rgb_feat, rd.permute(1, 0, 2), out_mask.t()[..., None])
        # ---- eval_func (src/model.py:1140-1160)
        vf = valid.float()[:, None]
        rad = o[:, 1:2]
        if noise is not None:
            rad = rad + noise.to(self.dev).reshape(-1, 1) * self.rand_noise_std
        rgba = torch.cat([vf * F.relu(rad), vf * o[:, 0:1] + (1.0 - vf) * 0.001, rgb], 1)
        return rgba, valid, geo

    def _ibr_head(self, rgb_feats, ray_diffs, mask):
        m = "mlp_tex."
        elu = F.elu
        V = rgb_feats.shape[1]
        d = elu(self._lin(elu(self._lin(ray_diffs, m + "ray_encoder.0")), m + "ray_encoder.2"))
        src_rgb = rgb_feats[..., :3]
        f = rgb_feats + d
        dot = ray_diffs[..., 3:4]
        e = torch.exp(torch.abs(self.P("mlp_tex.ani_al")) * (dot - 1))
        wgt = (e - torch.min(e, dim=1, keepdim=True)[0]) * mask
        wgt = wgt / (torch.sum(wgt, dim=1, keepdim=True) + 1e-8)
        mean = torch.sum(f * wgt, dim=1, keepdim=True)
        var = torch.sum(wgt * (f - mean) ** 2, dim=1, keepdim=True)
        x = torch.cat([mean.expand(-1, V, -1), var.expand(-1, V, -1), f], -1)
        x = elu(self._lin(elu(self._lin(x, m + "base_layer.0")), m + "base_layer.2"))
        pv = elu(self._lin(elu(self._lin(x * wgt, m + "vis_layer1.0")), m + "vis_layer1.2"))
        res, vis = pv[..., :-1], pv[..., -1:]
        x = x + res
        vis = torch.sigmoid(self._lin(elu(self._lin(x * torch.sigmoid(vis) * mask, m + "vis_layer2.0")), m + "vis_layer2.2")) * mask
        s = self._lin(elu(self._lin(elu(self._lin(torch.cat([x, vis, ray_diffs], -1), m + "out_layer.0")), m + "out_layer.2")), m + "out_layer.4")
        s = s.masked_fill(mask == 0, -1e4)
        return torch.sum(src_rgb * torch.softmax(s, dim=1), dim=1)

    # ---- one ray batch: coarse pass, importance sampling, fine pass (training flavour of batch_render_pifu_nerf)
    def render(self, cam_tar: Dict, bounds, pix_xy: torch.Tensor, rand: Optional[TrainRandom] = None, training=True, uniform=False,
               fine=True, S_c=64, S_f=64, znear=None, zfar=None) -> Dict[str, torch.Tensor]:
        r = self.renderer
        V = self.V
        tar = r.make_target(cam_tar, bounds, znear, zfar)
        R = pix_xy.shape[0]
        rand = rand or TrainRandom()
        lin = torch.linspace(0.0, 1.0, steps=S_c)
        t = None
        if not uniform:                                                          # stratified jitter (src/model.py:1226-1230)
            zt = lin[None, None, :].expand(1, R, -1)
            z_mid = 0.5 * (zt[..., 1:] + zt[..., :-1])
            z_lower, z_upper = torch.cat([zt[..., :1], z_mid], -1), torch.cat([z_mid, zt[..., -1:]], -1)
            t = (z_lower + rand.rand(1, R, S_c) * (z_upper - z_lower))[0].contiguous()
        rays, z = r.sample_rays(tar, pix_xy, S_c, t)
        beta = torch.clamp(self.P("sigmoid_beta"), min=2e-3)
        out = {}

        def one_pass(zz, S):
            N = R * S
            drop = rand.view_dropout(V) if (training and V > 1) else None
            noise = rand.randn(1, N, 1).reshape(-1) if (training and self.rand_noise_std > 0) else None
            rgba, valid, geo = self.shade(tar, rays, zz, drop, noise)
            color, depth, alpha, sdf, contrib = CompositeFn.apply(r, rgba.view(R, S, 5), zz, geo["sdf"].view(R, S), beta)
            return color, depth, alpha, sdf, contrib, rgba

        color, depth, alpha, _sdf, contrib, rgba_c = one_pass(z, S_c)
        out.update(tex_fg=color, depth=depth, alpha=alpha, z=z, rgba=rgba_c)
        if fine:
            u = r.linspace(S_f) if uniform else rand.rand(1, R, S_f)[0].contiguous()
            with torch.no_grad():
                z_f, z2 = r.importance(contrib.detach(), z, S_f, u)
            color2, depth2, alpha2, sdf2, _c2, rgba_f = one_pass(z2, S_c + S_f)
            out.update(tex_fg_fine=color2, depth_fine=depth2, alpha_fine=alpha2, sdf=sdf2, z_fine=z2, rgba_fine=rgba_f)
        return out


def patch_pixels(msk: torch.Tensor, width: int, height: int, out_h: int, out_w: int, rand: TrainRandom) -> torch.Tensor:
    """Random training patch (src/model.py:1172-1189): an out_h x out_w window centred on a random foreground pixel of `msk`
    (H,W), clamped to [0, min(width - 1, height - 1)] like the reference.  Returns (out_h * out_w, 2) int64 [x, y]."""
    coords = torch.stack(torch.where(msk)[::-1], -1)
    center = torch.zeros(1, 2, dtype=torch.long) if coords.shape[0] <= 0 else coords[rand.patch_center_index(coords.shape[0])][None].cpu()
    ys, xs = torch.meshgrid(torch.arange(0, out_h), torch.arange(0, out_w), indexing="ij")
    grids = torch.stack([xs, ys], -1).view(-1, 2) + (center - out_h // 2)
    return grids.clamp(0, min(width - 1, height - 1))


def allreduce_gradients(params, world: int, group=None):
    """DDP-style gradient averaging (train.py:58,65 of the reference runs Lightning DDP): ONE flat fp32 bucket of all
    render-path gradients, one all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests), unflatten."""
    import torch.distributed as dist
    ps = [p for p in params if p.grad is not None]
    if not ps or world == 1:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for p in ps:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()


def training_step(path: TrainableRenderPath, frame: Dict, pix_xy: torch.Tensor, target_rgb: torch.Tensor, optimizer=None, rand=None,
                  world: int = 1, lambda_c: float = 1.0, lambda_f: float = 10.0, **cfg):
    """Config E step: forward (coarse + fine), L1 colour losses (coarse x 1, fine x 10: configs/vanerf.json), backward through the
    path, gradient all-reduce over the ranks, optimiser step.  Returns the loss and the outputs."""
    if optimizer is not None:
        optimizer.zero_grad(set_to_none=True)
    path._tables = None
    out = path.render(frame["cam_tar"], frame["bounds"], pix_xy, rand=rand, **cfg)
    tgt = target_rgb.to(path.dev)
    loss = lambda_c * (out["tex_fg"] - tgt).abs().mean()
    if "tex_fg_fine" in out:
        loss = loss + lambda_f * (out["tex_fg_fine"] - tgt).abs().mean()
    with fp32_exact():
        loss.backward()
    allreduce_gradients(path.parameters(), world)
    if optimizer is not None:
        optimizer.step()
    return loss.detach(), out
