"""Host side of the B200 render path: owns a `vanerf_ctx`, packs weights, runs the per-frame setup and launches the
kernels through the C ABI (include/vanerf_b200.h).  torch is used for device memory, streams and the few per-frame
library ops the reference also leaves to torch (3x3 / 4x4 inverses, the TexVisFusion global-feature convolutions,
src/networks.py:273-279); everything per-ray / per-sample runs in libvanerf_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib as L
from . import ops as O
from . import weights as W

NUM_VERT = 1558


def _adaptive_avg_pool3(x):
    """adaptive_avg_pool2d(x, 3) as nine region means.  Same regions as torch (rows floor(i H / 3) .. ceil((i + 1) H / 3));
    torch's own CUDA kernel parallelises over output elements only and takes ~1 ms for a 512 x 334 map (two calls were
    half of the per-frame setup); nine block reductions take ~0.1 ms.  Summation order differs (~1e-7 relative)."""
    H, W = x.shape[-2:]
    out = x.new_empty(*x.shape[:-2], 3, 3)
    for i in range(3):
        h0, h1 = (i * H) // 3, -((-(i + 1) * H) // 3)
        for j in range(3):
            w0, w1 = (j * W) // 3, -((-(j + 1) * W) // 3)
            out[..., i, j] = x[..., h0:h1, w0:w1].mean((-2, -1))
    return out


def _np32(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().float().cpu().numpy()
    return np.ascontiguousarray(t, dtype=np.float32)


class Renderer:
    def __init__(self, device="cuda:0", lib: Optional[L.Lib] = None):
        self.lib = lib or L.get_lib()
        self.device = torch.device(device)
        if not self.lib.emulated and self.device.type != "cuda":
            raise L.VanerfError("vanerf_b200 runs on CUDA devices only (no CPU fallback)")
        self.ctx = C.c_void_p()
        dev_index = self.device.index or 0
        self.lib.check(None, self.lib.dll.vanerf_ctx_create(C.byref(self.ctx), dev_index), "vanerf_ctx_create")
        self.sd: Optional[Dict[str, torch.Tensor]] = None
        self.frame = None
        self._keep = []          # host arrays referenced by C structs during a call
        self._tabs = {}
        self.handle = O.register_renderer(self)       # what the torch custom ops (vanerf_b200/ops.py) take as `ctx`

    def __del__(self):
        try:
            if self.ctx:
                self.lib.dll.vanerf_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------ plumbing
    @property
    def stream(self):
        if self.lib.emulated:
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _ptr(self, t: Optional[torch.Tensor]):
        if t is None:
            return None
        assert t.is_contiguous()
        if not self.lib.emulated:
            assert t.is_cuda, "device tensor expected"
        return C.c_void_p(t.data_ptr())

    def empty(self, shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    @property
    def launches(self) -> int:
        return int(self.lib.dll.vanerf_launch_count(self.ctx))

    def linspace(self, n: int) -> torch.Tensor:
        """torch.linspace(0,1,n) computed by torch on the host like the reference (src/model.py:1222,1440)."""
        if n not in self._tabs:
            self._tabs[n] = torch.linspace(0.0, 1.0, steps=n).to(self.device)
        return self._tabs[n]

    # ------------------------------------------------------------------------------------------ weights
    def load_state_dict(self, state_dict):
        """Accepts the reference `VANeRF.state_dict()` / a Lightning checkpoint's `state_dict` (numpy or torch)."""
        sd = W.strip_prefix(state_dict)
        folded = W.fold(sd)
        keep = []

        def lin(tag):
            w = np.ascontiguousarray(folded[tag + ".w"], np.float32)
            b = folded.get(tag + ".b")
            keep.append(w)
            v = L.VLinear()
            v.w = w.ctypes.data
            v.out_dim, v.in_dim = w.shape
            if b is not None:
                b = np.ascontiguousarray(b, np.float32)
                keep.append(b)
                v.b = b.ctypes.data
            return v
        vw = L.VWeights()
        for name, tags in [("geo_at", ["geo_at0", "geo_at1"]), ("geo_f", ["geo_f0", "geo_f1"]),
                           ("geo8_at", ["geo8_at0", "geo8_at1"]), ("geo8_f", ["geo8_f0", "geo8_f1"]),
                           ("mlp", ["mlp0", "mlp1", "mlp2", "mlp3"]), ("post", ["post0", "post1", "post2"]),
                           ("tex_at", ["tex_at0", "tex_at1"]), ("tex_f", ["tex_f0", "tex_f1"]),
                           ("ray", ["ray0", "ray1"]), ("base", ["base0", "base1"]), ("vis1", ["vis10", "vis11"]),
                           ("vis2", ["vis20", "vis21"]), ("outl", ["outl0", "outl1", "outl2"])]:
            arr = getattr(vw, name)
            for i, t in enumerate(tags):
                arr[i] = lin(t)
        vw.compress = lin("compress")
        vw.ani_al = float(folded["ani_al"][0])
        vw.sigmoid_beta = float(folded["sigmoid_beta"][0])
        self.lib.check(self.ctx, self.lib.dll.vanerf_load_weights(self.ctx, C.byref(vw), self.stream), "vanerf_load_weights")
        # per-frame TexVisFusion convolution stacks stay in torch (cuDNN), on the device
        self.sd = {k: torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).float().to(self.device).contiguous()
                   for k, v in sd.items() if k.startswith(("tex_vis_fusion.fconv3", "tex_vis_fusion.fconv4", "tex_vis_fusion.fconv_gt"))}
        self.sigmoid_beta = max(2e-3, float(folded["sigmoid_beta"][0]))

    # ------------------------------------------------------------------------------------------ per frame
    @torch.no_grad()
    def global_vertex_feature(self, img, feat_tex):
        """TexVisFusion global feature per vertex, (V,1558,18) (src/networks.py:273-279).  LayerNorm shapes follow
        the actual map sizes."""
        sd = self.sd
        if not self.lib.emulated and img.is_cuda and not torch.is_grad_enabled():
            # product path: the stacks as kernels of the library (csrc/gfeat.cuh); torch below only serves the host-emulation
            # tests and the differentiable training graph
            V, _, H, Wd = img.shape
            img, feat_tex = img.contiguous(), feat_tex.contiguous()
            w = L.VGfeatWeights()
            for stack, pre in ((w.img, "tex_vis_fusion.fconv4"), (w.tex, "tex_vis_fusion.fconv3"), (w.gt, "tex_vis_fusion.fconv_gt")):
                for field, key in (("conv0", ".0.weight"), ("ln1_w", ".1.weight"), ("ln1_b", ".1.bias"), ("conv3", ".3.weight"),
                                   ("ln4_w", ".4.weight"), ("ln4_b", ".4.bias")):
                    t = sd[pre + key]
                    assert t.is_contiguous() and t.dtype == torch.float32
                    setattr(stack, field, t.data_ptr())
            assert tuple(sd["tex_vis_fusion.fconv4.1.weight"].shape) == (H, Wd) and tuple(sd["tex_vis_fusion.fconv3.1.weight"].shape) == tuple(feat_tex.shape[-2:]), \
                "LayerNorm maps of the per-frame stacks must match the image / texture map sizes"
            out = self.empty((V, NUM_VERT, 18))
            st = self.lib.dll.vanerf_global_vertex_feature(self.ctx, C.byref(w), self._ptr(img), self._ptr(feat_tex), V, H, Wd, feat_tex.shape[2],
                                                           feat_tex.shape[3], self._ptr(out), self.stream)
            self.lib.check(self.ctx, st, "vanerf_global_vertex_feature")
            return out
        # cuDNN would take TF32 for these convolutions by default (the reference's own GPU run does);
        # the CPU oracle is the arbiter, so the per-frame stacks run in true fp32.
        with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True, allow_tf32=False):
            return self._global_vertex_feature(img, feat_tex, sd)

    @staticmethod
    def _global_vertex_feature(img, feat_tex, sd):
        def stack(x, pre):
            x = F.conv2d(x, sd[pre + ".0.weight"], padding=1)
            x = F.relu(F.layer_norm(x, x.shape[-2:], sd[pre + ".1.weight"], sd[pre + ".1.bias"], 1e-6))
            x = F.conv2d(x, sd[pre + ".3.weight"], padding=1)
            x = F.relu(F.layer_norm(x, x.shape[-2:], sd[pre + ".4.weight"], sd[pre + ".4.bias"], 1e-6))
            return _adaptive_avg_pool3(x)
        gf = stack(feat_tex, "tex_vis_fusion.fconv3")
        gi = stack(img, "tex_vis_fusion.fconv4")
        g = torch.cat([gi.reshape(*gi.shape[:2], -1), gf.reshape(*gf.shape[:2], -1)], -1)
        p = "tex_vis_fusion.fconv_gt"
        x = F.conv1d(g, sd[p + ".0.weight"], padding=1)
        x = F.relu(F.layer_norm(x, (18,), sd[p + ".1.weight"], sd[p + ".1.bias"], 1e-6))
        x = F.conv1d(x, sd[p + ".3.weight"], padding=1)
        x = F.relu(F.layer_norm(x, (18,), sd[p + ".4.weight"], sd[p + ".4.bias"], 1e-6))
        return x.contiguous()

    @torch.no_grad()
    def set_frame(self, img, cam_in, targets, sp_data, feat_geo, feat_tex, src_foreground_mask):
        """Per-frame setup from the reference-layout inputs (B = 1): img (V,3,H,W), cam_in dict (decode_batch,
        src/model.py:313-317), targets{vert_world,face_world}, sp_data{extrin,kpt3d}, feat_geo [g0,g1], feat_tex,
        src_foreground_mask (1,V,1,H,W) bool."""
        assert self.sd is not None, "load_state_dict first"
        V = img.shape[0]
        H, Wd = int(cam_in["height"]), int(cam_in["width"])
        dev = self.device
        f32 = lambda t: t.detach().to(dev, torch.float32).contiguous()
        img_d, g0, g1, tx = f32(img), f32(feat_geo[0]), f32(feat_geo[1]), f32(feat_tex)
        fg = src_foreground_mask.reshape(V, H, Wd).to(dev).to(torch.uint8).contiguous()
        gfeat = self.global_vertex_feature(img_d, tx)
        KRT = cam_in["KRT"].detach().float()
        src_pos = torch.inverse(KRT)[:, :3, 3]                        # src/model.py:937-938
        znear, zfar = float(cam_in["znear"]), float(cam_in["zfar"])
        verts = _np32(targets["vert_world"]).reshape(-1, 3)
        faces = np.ascontiguousarray(targets["face_world"].detach().cpu().numpy().reshape(-1, 3).astype(np.int32))
        host = dict(KRT=_np32(KRT), extrin=_np32(sp_data["extrin"]), src_pos=_np32(src_pos),
                    kpt=_np32(sp_data["kpt3d"]).reshape(-1, 3), verts=verts, faces=faces)
        fr = L.VFrame()
        fr.n_views, fr.height, fr.width = V, H, Wd
        fr.znear, fr.zfar, fr.z_range = znear, zfar, float(np.float32(zfar - znear))
        fr.KRT, fr.extrin = host["KRT"].ctypes.data, host["extrin"].ctypes.data
        fr.src_cam_pos, fr.kpt3d = host["src_pos"].ctypes.data, host["kpt"].ctypes.data
        fr.verts, fr.faces = verts.ctypes.data, faces.ctypes.data
        fr.n_verts, fr.n_faces = verts.shape[0], faces.shape[0]
        fr.img, fr.fg_mask = img_d.data_ptr(), fg.data_ptr()
        fr.feat_geo0, fr.g0_h, fr.g0_w = g0.data_ptr(), g0.shape[2], g0.shape[3]
        fr.feat_geo1, fr.g1_h, fr.g1_w = g1.data_ptr(), g1.shape[2], g1.shape[3]
        fr.feat_tex, fr.t_h, fr.t_w = tx.data_ptr(), tx.shape[2], tx.shape[3]
        fr.vert_gfeat = gfeat.data_ptr()
        vert_vis = self.empty((V, verts.shape[0]))
        self.lib.check(self.ctx, self.lib.dll.vanerf_frame_setup(self.ctx, C.byref(fr), self._ptr(vert_vis), self.stream),
                       "vanerf_frame_setup")
        self.frame = dict(V=V, H=H, W=Wd, znear=znear, zfar=zfar, vert_vis=vert_vis, gfeat=gfeat,
                          inputs=(img_d, g0, g1, tx, fg))
        return vert_vis

    def make_target(self, cam_tar, bounds, znear=None, zfar=None) -> L.VTarget:
        """Per-render 3x3 matrices by the same torch calls as the reference (src/model.py:1208-1213)."""
        K, RT = cam_tar["K"].detach().float().cpu(), cam_tar["RT"].detach().float().cpu()
        inv_K = torch.inverse(K[:, :3, :3]).transpose(1, 2)[0].contiguous().numpy()
        R = RT[0, :3, :3].contiguous().numpy()
        cam_pos = (-torch.bmm(RT[:, :3, 3][:, None], RT[:, :3, :3]))[0, 0].numpy()
        t = L.VTarget()
        t.inv_K[:] = inv_K.reshape(-1).tolist()
        t.R[:] = R.reshape(-1).tolist()
        t.cam_pos[:] = cam_pos.tolist()
        t.znear = float(cam_tar.get("znear", self.frame["znear"]) if znear is None else znear)
        t.zfar = float(cam_tar.get("zfar", self.frame["zfar"]) if zfar is None else zfar)
        t.bounds[:] = _np32(bounds).reshape(-1).tolist()
        return t

    # ------------------------------------------------------------------------------------------ stages
    # Every stage goes through the torch custom ops of vanerf_b200/ops.py (torch.ops.vanerf_b200.*), which bind the C ABI.
    def sample_rays(self, tar, pix_xy: torch.Tensor, n_samples: int, t: Optional[torch.Tensor] = None):
        """t: optional (R, n_samples) per-ray interpolation parameters (training: stratified jitter, src/model.py:1226-1230);
        default = linspace(0, 1, n_samples) shared by all rays (uniform=True)."""
        pix = pix_xy.to(self.device, torch.int32).contiguous()
        ztab = self.linspace(n_samples) if t is None else t.to(self.device, torch.float32).contiguous()
        return torch.ops.vanerf_b200.sample_rays(self.handle, O.pack_target(tar), pix, ztab)

    def geom_query(self, tar, rays, z, want_pts=True):
        pts, sdf, face, nn, qvis = torch.ops.vanerf_b200.geom_query(self.handle, O.pack_target(tar), rays, z)
        return dict(pts=pts if want_pts else None, sdf=sdf, face=face, nn=nn, qvis=qvis)

    def shade(self, tar, rays, z, geo, precision=L.FP32, want_raw=True, want_latent=False):
        R, S = z.shape
        N = R * S
        if want_latent:                 # test hook (vanerf_shade_debug*): also returns MLPUNetFusion's pooled latent
            rgba, valid = self.empty((N, 5)), self.empty((N,), torch.uint8)
            raw = self.empty((N, 5)) if want_raw else None
            lat = self.empty((N, 128))
            fn = self.lib.dll.vanerf_shade_debug if precision == L.FP32 else self.lib.dll.vanerf_shade_debug_bf16
            st = fn(self.ctx, C.byref(tar), self._ptr(rays), self._ptr(z), R, S, self._ptr(geo["sdf"]),
                                                 self._ptr(geo["nn"]), self._ptr(geo["qvis"]), self._ptr(rgba), self._ptr(valid),
                                                 self._ptr(raw), self._ptr(lat), self.stream)
            self.lib.check(self.ctx, st, "vanerf_shade_debug")
            return rgba, valid, raw, lat
        rgba, valid, raw = torch.ops.vanerf_b200.shade(self.handle, O.pack_target(tar), rays, z, geo["sdf"], geo["nn"], geo["qvis"], int(precision))
        return rgba, valid, (raw if want_raw else None)

    def composite(self, rgba, z, mesh_sdf):
        color, depth, alpha, sdf, contrib = torch.ops.vanerf_b200.composite(self.handle, rgba, z, mesh_sdf)
        return dict(color=color, depth=depth, alpha=alpha, sdf=sdf, contrib=contrib)

    def importance(self, contrib, z, n_fine: int, u: Optional[torch.Tensor] = None):
        u = self.linspace(n_fine) if u is None else u.to(self.device, torch.float32).contiguous()
        return torch.ops.vanerf_b200.importance_sample(self.handle, contrib, z, u)

    def query_points(self, tar, pts, view, query_sdf=None, query_vis=None, precision=L.FP32):
        """VANeRF.query on explicit points: pts, view (N,3).  Returns raw (N,5) = [o0,o1,r,g,b], valid (N,), rgba."""
        pts = pts.to(self.device, torch.float32).contiguous()
        view = view.to(self.device, torch.float32).contiguous()
        sdf = self.empty((0,)) if query_sdf is None else query_sdf.to(self.device, torch.float32).contiguous()
        qv = self.empty((0,), torch.uint8) if query_vis is None else query_vis.to(self.device).to(torch.uint8).contiguous()
        return torch.ops.vanerf_b200.query_points(self.handle, O.pack_target(tar), pts, view, sdf, qv, int(precision))

    def finish(self):
        """Completion check (synchronises the stream): raises if a tensor-core launch gave up on a bounded wait since the
        last check, i.e. if results of the bf16 path are invalid (include/vanerf_b200.h: vanerf_tc_check)."""
        self.lib.check(self.ctx, self.lib.dll.vanerf_tc_check(self.ctx, self.stream), "vanerf_tc_check")

    def tc_error(self) -> int:
        """Nonzero when a bounded wait inside a tensor-core kernel gave up (synchronises the device first)."""
        if not self.lib.emulated:
            torch.cuda.synchronize(self.device)
        return int(self.lib.dll.vanerf_tc_error(self.ctx))

    def tc_selftest(self, A: torch.Tensor, W: torch.Tensor) -> torch.Tensor:
        """bf16(A (128,K)) @ bf16(W (N,K))^T through one tcgen05 step (test hook)."""
        K, N = A.shape[1], W.shape[0]
        A = A.to(self.device, torch.float32).contiguous()
        Wh = np.ascontiguousarray(W.detach().cpu().numpy(), np.float32)
        D = self.empty((128, (N + 15) // 16 * 16))
        st = self.lib.dll.vanerf_tc_selftest(self.ctx, self._ptr(A), Wh.ctypes.data, K, N, self._ptr(D), self.stream)
        self.lib.check(self.ctx, st, "vanerf_tc_selftest")
        return D[:, :N]

    KERNEL_CLASSES = ("setup", "rays", "geom", "gather", "mlp", "composite", "importance")

    def timing(self, on: bool):
        self.lib.check(self.ctx, self.lib.dll.vanerf_timing_enable(self.ctx, int(on)), "vanerf_timing_enable")

    def timing_read(self, reset=True):
        ms = (C.c_double * 7)()
        cnt = (C.c_int64 * 7)()
        self.lib.check(self.ctx, self.lib.dll.vanerf_timing_read(self.ctx, ms, cnt, int(reset)), "vanerf_timing_read")
        return {k: (ms[i], cnt[i]) for i, k in enumerate(self.KERNEL_CLASSES)}

    def set_reuse_coarse(self, on: bool):
        """Fine pass evaluates only the new depths and reuses the coarse pass for the rest (identical output bits, a third
        fewer network evaluations per ray at 64 + 64; include/vanerf_b200.h: vanerf_set_reuse_coarse).  Default off."""
        self.lib.check(self.ctx, self.lib.dll.vanerf_set_reuse_coarse(self.ctx, int(bool(on))), "vanerf_set_reuse_coarse")

    def set_reuse_geometry(self, on: bool):
        """Fine pass queries the mesh for the new depths only (default on; identical bits, the networks still evaluate every
        merged sample; include/vanerf_b200.h: vanerf_set_reuse_geometry)."""
        self.lib.check(self.ctx, self.lib.dll.vanerf_set_reuse_geometry(self.ctx, int(bool(on))), "vanerf_set_reuse_geometry")

    def render_rays(self, tar, pix_xy, n_coarse=64, n_fine=64, fine=True, precision=L.FP32):
        """One call for a ray batch: coarse pass, importance sampling, fine pass (src/model.py:1103-1360).
        Returns (R,8) rows [r,g,b,depth,alpha,sdf,0,0] for the coarse and the fine pass."""
        pix = pix_xy.to(self.device, torch.int32).contiguous()
        oc, of = torch.ops.vanerf_b200.render_rays(self.handle, O.pack_target(tar), pix, self.linspace(n_coarse), self.linspace(n_fine),
                                                   bool(fine), int(precision))
        return oc, (of if fine else None)
