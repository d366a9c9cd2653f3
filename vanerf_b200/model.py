"""Drop-in mirror of the reference's model/renderer call surface for the render path.

`VANeRF` keeps the names, argument meaning and return layouts of the reference class (src/model.py:604-1570) —
`render_pifu_nerf`, `batch_render_pifu_nerf`, `query`, `rgba2out`, `importance_sample`, `ray_bbox_intersection`,
`sdf_activation`, `attach_*_feat` — so `VANeRFLightningModule.render_full_nerf_image` / `render_novel_views`
(src/model.py:488-545) can run on top of it unmodified.  Everything per ray / per sample is executed by
libvanerf_b200.so through `Renderer`; the CNN encoders are the step in front of the path (cuDNN through
vanerf_b200/encoders.py): their feature maps are passed in (`feat_geo`, `feat_tex`), attached with `attach_im_feat(feat_geo=...,
feat_tex=...)`, or computed from the source images once `build_encoders()` / a checkpoint with encoder weights has provided them.

Supported configuration: batch size 1 (like every reference config), 1..4 source views (3 on the bf16 path;
V-generalisation of SURVEY.md Appendix C), any H x W.  Inference (`uniform=True`, eval mode) runs the fused kernels.  The
training branch (`train()`, or `uniform=False`: random patch, stratified jitter, density noise, view dropout, random importance
samples; src/model.py:804-810,1155-1156,1172-1189,1226-1230,1439-1442) runs the differentiable graph of vanerf_b200/train.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib as L
from .renderer import Renderer

PRECISIONS = {"fp32": L.FP32, "bf16": L.BF16}


class VANeRF:
    def __init__(self, cfg: Optional[dict] = None, device="cuda:0", precision="fp32", lib: Optional[L.Lib] = None):
        self.kwargs = (cfg or {}).get("models", {}).get("VANeRF", {}) if cfg else {}
        self.dr_level = self.kwargs.get("dr_level", 5)
        self.renderer = Renderer(device, lib)
        self.device = self.renderer.device
        self.precision = PRECISIONS[precision]
        self.training = False
        self.feat_geo = None
        self.feat_tex = None
        self.encoders = None                  # vanerf_b200.encoders.Encoders (build_encoders / checkpoint with encoder weights)
        self._frame_key = None
        self._frame_refs = None
        self._state_dict = None
        self.train_path = None
        self.train_out_h = self.train_out_w = int(self.kwargs.get("train_out_h", 64))       # configs/vanerf.json:46-47

    # ---------------------------------------------------------------- module-like surface
    def train(self, mode=True):
        """Training mode (src/model.py `net.training` paths): `batch_render_pifu_nerf` then runs the differentiable, unfused graph
        of vanerf_b200/train.py (random patch, stratified jitter, view dropout, density noise, random importance samples; fp32)
        on `self.train_path`, whose parameters (reference state_dict keys) an optimiser can update.  `eval()` copies the updated
        parameters back into the packed weights of the fused inference kernels."""
        if mode:
            if self._state_dict is None:
                raise L.VanerfError("load_state_dict before train()")
            if self.train_path is None:
                from .train import TrainableRenderPath
                self.train_path = TrainableRenderPath(self._state_dict, self.device, self.renderer.lib,
                                                      rand_noise_std=float(self.kwargs.get("rand_noise_std", 0.01)))
        self.training = bool(mode)
        return self

    def eval(self):
        if self.training and self.train_path is not None:
            self.load_state_dict(self.train_path.state_dict_ref())
        self.training = False
        return self

    def parameters(self):
        """Render-path parameters of the training graph (after train())."""
        return self.train_path.parameters() if self.train_path is not None else iter(())

    def load_state_dict(self, state_dict, strict=False):
        self.renderer.load_state_dict(state_dict)
        self._state_dict = dict(state_dict)
        enc = {k: v for k, v in state_dict.items() if k.startswith(("geo_encoder.", "tex_encoder."))}
        if enc:                               # a full reference checkpoint: the encoders come with it
            self.build_encoders()
            self.encoders.load_state_dict({k: torch.as_tensor(v) for k, v in enc.items()}, strict=True)
        self._frame_key = None
        return self

    def build_encoders(self, seed: Optional[int] = None, autocast_dtype=None):
        """Creates the CNN encoders of the step in front of the path (vanerf_b200/encoders.py; reference `geo_encoder`,
        `tex_encoder`, src/model.py:630-654).  `load_state_dict` calls this when the checkpoint has `geo_encoder.*` entries;
        `seed` fills them with name-keyed synthetic weights instead (no checkpoint is available offline)."""
        from .encoders import Encoders, seeded_state_dict
        m = self.kwargs
        self.encoders = Encoders(m.get("geo_args"), m.get("tex_args"), int(m.get("ds_geo", 1)), int(m.get("ds_tex", 1))).eval()
        if seed is not None:
            self.encoders.geo_encoder.load_state_dict(seeded_state_dict(self.encoders.geo_encoder, seed), strict=True)
            self.encoders.tex_encoder.load_state_dict(seeded_state_dict(self.encoders.tex_encoder, seed), strict=True)
        self.encoders.to(self.device)
        self.encoders.autocast_dtype = autocast_dtype
        return self.encoders

    def attach_im_feat(self, im=None, return_val=False, feat_geo=None, feat_tex=None):
        """src/model.py:700-709.  The maps are either supplied by the caller (`feat_geo=[g0, g1]`, `feat_tex=`) or computed from
        the source images `im` (V,3,H,W in [0,1]) by the encoders (`build_encoders` / a checkpoint with encoder weights)."""
        if feat_geo is None:
            feat_geo = self.attach_geo_feat(im, return_val=True)
        if feat_tex is None:
            feat_tex = self.attach_tex_feat(im, return_val=True)
        self.feat_geo, self.feat_tex = feat_geo, feat_tex
        if return_val:
            return {"feat_geo": feat_geo, "feat_tex": feat_tex}

    def attach_geo_feat(self, im, return_val=False):
        """src/model.py:711-724: average-pool `ds_geo` times, 2 x - 1, HGFilterV2 -> [geo0 (V,64,H/8,W/8), geo1 (V,8,H/2,W/2)]."""
        if self.encoders is not None and im is not None:
            self.feat_geo = self.encoders.encode_geo(im.to(self.device))
        elif self.feat_geo is None:
            raise L.VanerfError("no geometry feature maps: build_encoders() / load a checkpoint with geo_encoder.* weights and pass the "
                                "source images, or attach_im_feat(feat_geo=[g0, g1], feat_tex=...)")
        return self.feat_geo if return_val else None

    def attach_tex_feat(self, im, return_val=False):
        """src/model.py:726-738: average-pool `ds_tex` times, 2 x - 1, ResBlkEncoder -> (V,8,H/4,W/4)."""
        if self.encoders is not None and im is not None:
            self.feat_tex = self.encoders.encode_tex(im.to(self.device))
        elif self.feat_tex is None:
            raise L.VanerfError("no texture feature map: build_encoders() / load a checkpoint with tex_encoder.* weights and pass the "
                                "source images, or attach_im_feat(feat_geo=[g0, g1], feat_tex=...)")
        return self.feat_tex if return_val else None

    def detach_im_feat(self):
        self.feat_geo = self.feat_tex = None

    def sdf_activation(self, input):
        """src/model.py:879-882."""
        beta = self.renderer.sigmoid_beta
        return torch.sigmoid(input / beta) / beta

    # ---------------------------------------------------------------- per-frame state
    def _ensure_frame(self, img_in, cam_in, targets, sp_data, feat_geo, feat_tex, fg_mask):
        """Runs the per-frame setup unless EVERY per-frame input is the very tensor object (same identity, same in-place
        version counter) the current frame was built from.  The tensors of the current frame are kept referenced, so their
        addresses and ids cannot be recycled by a later batch of the same shapes; an in-place update of any of them bumps
        `_version` and triggers a new setup.  (Keying on `data_ptr()` is wrong: the caching allocator hands a freed block to
        the next frame's tensors.)"""
        tensors = (img_in, feat_geo[0], feat_geo[1], feat_tex, fg_mask, targets["vert_world"], targets["face_world"],
                   cam_in["KRT"], sp_data["extrin"], sp_data["kpt3d"])
        scalars = (int(cam_in["height"]), int(cam_in["width"]), float(cam_in["znear"]), float(cam_in["zfar"]))
        key = tuple((id(t), t._version) for t in tensors) + scalars
        if key != self._frame_key:
            if img_in.shape[0] > L.MAX_VIEWS_BF16 and self.precision == L.BF16:
                raise L.VanerfError(f"the bf16 tensor-core path supports up to {L.MAX_VIEWS_BF16} source views "
                                    f"(got {img_in.shape[0]}); use precision='fp32' (up to {L.MAX_VIEWS})")
            self.vert_vis = self.renderer.set_frame(img_in, cam_in, targets, sp_data, feat_geo, feat_tex, fg_mask)
            self._frame_key = key
            self._frame_refs = tensors
        return self.vert_vis

    def invalidate_frame(self):
        """Forces the next call to run the per-frame setup again."""
        self._frame_key = None
        self._frame_refs = None

    # ---------------------------------------------------------------- VANeRF.query (src/model.py:748-877)
    def query(self, pts, cam, hand_type, targets, feat_geo=None, feat_tex=None, vert=None, vert_vis=None,
              query_vis=None, query_sdf=None, closest_face=None, n_views=1, sp_data={}, tx_data={}, view=None,
              n_pts_samples=-1, **kwargs):
        """pts (1,N,3), view (1,N,3) -> out (1,N,5) = [o0,o1,r,g,b], valid (1,N,1) bool.  `query_sdf` (n_views,N) /
        `query_vis` (n_views,N,1) are honoured when given (the reference passes cal_vis_sdf_batch's results in);
        otherwise they are computed from the mesh.  Per-view vertex visibility is recomputed from the frame."""
        assert pts.shape[0] == 1, "batch size 1"
        feat_geo = feat_geo if feat_geo is not None else self.feat_geo
        feat_tex = feat_tex if feat_tex is not None else self.feat_tex
        self._ensure_frame(tx_data["img"], cam, targets, sp_data, feat_geo, feat_tex, kwargs["src_foreground_mask"])
        r = self.renderer
        bounds = kwargs.get("bounds", torch.zeros(1, 2, 3))
        tar = r.make_target({"K": cam["K"][:1], "RT": sp_data["extrin"][:1]}, bounds)   # only cam-independent fields are used
        sdf = query_sdf.reshape(-1, pts.shape[1])[0] if query_sdf is not None else None
        qv = query_vis.reshape(n_views, -1) if query_vis is not None else None
        raw, valid, _ = r.query_points(tar, pts[0], view[0], sdf, qv, self.precision)
        return raw[None], valid.bool()[None, :, None]

    # ---------------------------------------------------------------- static helpers kept from the reference
    @staticmethod
    def rgba2out(self, rgba, z, vert_sdf):
        """src/model.py:1465-1494.  rgba (1,R,S,5), z (1,R,S), vert_sdf (1,R,S,1) -> color, depth, alpha, contrib, sdf."""
        r = self.renderer
        c = r.composite(rgba[0].reshape(-1, 5).contiguous(), z[0].contiguous(), vert_sdf[0].reshape(z[0].shape).contiguous())
        return c["color"][None], c["depth"][None], c["alpha"][None], c["contrib"][None], c["sdf"][None]

    def importance_sample_merged(self, contrib, z, sample_per_ray, u=None):
        """importance_sample + sort(cat[z, z_fine]) in one kernel (src/model.py:1301-1307).  contrib, z: (1,R,S)."""
        zf, zall = self.renderer.importance(contrib[0].contiguous(), z[0].contiguous(), sample_per_ray, u)
        return zf[None], zall[None]

    def importance_sample(self, contrib, z, sample_per_ray, uniform=False):
        """src/model.py:1425-1462 with the reference's argument convention: contrib (1,R,S-2) = contrib[...,1:-1],
        z (1,R,S-1) = z_mid; uniform=False draws one row of uniform random numbers per ray (training)."""
        import ctypes as C
        r = self.renderer
        R, nb = contrib.shape[1], contrib.shape[2]
        D = nb + 2
        ci, zm = contrib[0].to(self.device).float().contiguous(), z[0].to(self.device).float().contiguous()
        zf = r.empty((R, sample_per_ray))
        # uniform=False: one row of uniform random numbers per ray (src/model.py:1442), drawn with torch like the reference
        u = r.linspace(sample_per_ray) if uniform else torch.rand(R, sample_per_ray).to(self.device).contiguous()
        st = r.lib.dll.vanerf_importance_mid(r.ctx, r._ptr(ci), r._ptr(zm), R, D, r._ptr(u), sample_per_ray, 0 if uniform else 1, r._ptr(zf), r.stream)
        r.lib.check(r.ctx, st, "vanerf_importance_mid")
        return zf[None]

    @staticmethod
    def ray_bbox_intersection(bounds, orig, direct, boffset=(-0.01, 0.01)):
        """src/model.py:1497-1570 semantics on torch tensors (host-side helper; the render path clips rays inside
        vanerf_sample_rays)."""
        b = bounds[0] + torch.tensor([boffset[0], boffset[1]], device=bounds.device)[:, None]
        d = direct[0].clone()
        d[d.abs() < 1e-5] = 1e-5
        o = orig[0].expand(d.shape[0], -1)
        t6 = ((b[None] - o[:, None]) / d[:, None]).reshape(-1, 6)
        p6 = t6[..., None] * d[:, None] + o[:, None]
        bf = b.reshape(-1)
        eps = 1e-6
        inside = ((p6[..., 0] >= bf[0] - eps) & (p6[..., 0] <= bf[3] + eps) & (p6[..., 1] >= bf[1] - eps) &
                  (p6[..., 1] <= bf[4] + eps) & (p6[..., 2] >= bf[2] - eps) & (p6[..., 2] <= bf[5] + eps))
        hit = inside.sum(-1) == 2
        near, far = torch.ones(d.shape[0], device=d.device), torch.ones(d.shape[0], device=d.device)
        if hit.any():
            pts = p6[hit][inside[hit]].reshape(-1, 2, 3)
            nr = torch.linalg.norm(d[hit], dim=1)
            d0 = torch.linalg.norm(pts[:, 0] - o[hit], dim=1) / nr
            d1 = torch.linalg.norm(pts[:, 1] - o[hit], dim=1) / nr
            near[hit], far[hit] = torch.minimum(d0, d1), torch.maximum(d0, d1)
        return near[None, :, None], far[None, :, None], hit[None, :, None]

    # ---------------------------------------------------------------- batch_render_pifu_nerf (src/model.py:1103-1422)
    @staticmethod
    def batch_render_pifu_nerf(net, img_in, cam_in, hand_type, targets, n_views, cam_tar, level=2, stride=0, tar_img=None,
                               feat_geo=None, feat_tex=None, mano_vert_world=None, sp_data={}, objcenter=None, **config):
        assert cam_tar["K"].shape[0] == 1, "batch size 1"
        S_c, S_f = config.get("sample_per_ray_c", 64), config.get("sample_per_ray_f", 64)
        fine = config.get("fine", False)
        feat_geo = feat_geo if feat_geo is not None else net.feat_geo
        feat_tex = feat_tex if feat_tex is not None else net.feat_tex
        if net.training or not config.get("uniform", False):
            return net._batch_render_train(img_in, cam_in, hand_type, targets, n_views, cam_tar, level, stride, tar_img, feat_geo, feat_tex,
                                           sp_data, **config)
        width, height = int(cam_tar.get("width", cam_in["width"])), int(cam_tar.get("height", cam_in["height"]))
        step = 2 ** (level - 1)
        assert width % step == 0 and height % step == 0
        if isinstance(stride, torch.Tensor):
            sx, sy = int(stride.reshape(-1, 2)[0, 0].item()), int(stride.reshape(-1, 2)[0, 1].item())
        else:
            sx = sy = int(stride)
        assert max(sx, sy) < step
        out_w, out_h = width // step, height // step
        dev = net.device
        ys, xs = torch.meshgrid(torch.arange(0, height, step, device=dev), torch.arange(0, width, step, device=dev), indexing="ij")
        grids = torch.stack([xs + sx, ys + sy], -1).reshape(-1, 2)
        if "pixel_override" in config:                      # explicit target pixels (same hook as the oracle harness)
            grids = config["pixel_override"].reshape(-1, 2).to(dev)
            out_h, out_w = 1, grids.shape[0]
        index = (grids[:, 0] + grids[:, 1] * width).long()
        vert_vis = net._ensure_frame(img_in, cam_in, targets, sp_data, feat_geo, feat_tex, config["src_foreground_mask"])
        r = net.renderer
        tar = r.make_target(cam_tar, config["bounds"], cam_tar.get("znear", cam_in["znear"]), cam_tar.get("zfar", cam_in["zfar"]))
        # reuse_coarse (extension, default False = the reference's evaluation count): identical output bits, the fine pass
        # evaluates only the S_f new depths (Renderer.set_reuse_coarse)
        r.set_reuse_coarse(bool(config.get("reuse_coarse", False)))
        oc, of = r.render_rays(tar, grids, S_c, S_f, fine, net.precision)
        img = lambda t, c: t.reshape(out_h, out_w, c).permute(2, 0, 1)[None]
        out = {"tex_fg": img(oc[:, :3], 3), "depth": oc[:, 3].reshape(1, out_h, out_w), "alpha": oc[:, 4].reshape(1, out_h, out_w)}
        if fine:
            out.update({"tex_fg_fine": img(of[:, :3], 3), "depth_fine": of[:, 3].reshape(1, out_h, out_w),
                        "alpha_fine": of[:, 4].reshape(1, out_h, out_w), "sdf": of[:, 5].reshape(1, out_h, out_w)})
        # auxiliary gathers at the ray pixels (src/model.py:1361-1418); vis_img (GAN supervision) is out of scope
        with torch.no_grad():
            if tar_img is not None:
                t = tar_img.reshape(*tar_img.shape[:2], -1).to(dev)
                out["tar_img"] = torch.gather(t, 2, index[None, None].expand(t.shape[0], 3, -1)).view(t.shape[0], 3, out_h, out_w)
            m = config["src_foreground_mask"].reshape(1, -1, height * width)[:, :1].to(dev)
            out["input_mask"] = torch.gather(m, 2, index[None, None]).view(1, 1, out_h, out_w)
            im = img_in.to(dev)[::n_views].reshape(1, 3, -1)
            out["img_in"] = torch.gather(im, 2, index[None, None].expand(-1, 3, -1)).view(1, 3, out_h, out_w)
        out["vert_vis"] = vert_vis[:, :, None]
        return out

    def _batch_render_train(self, img_in, cam_in, hand_type, targets, n_views, cam_tar, level, stride, tar_img, feat_geo, feat_tex, sp_data,
                            **config):
        """Training flavour of batch_render_pifu_nerf (src/model.py:1103-1422 with net.training / uniform=False): differentiable
        outputs from vanerf_b200.train.TrainableRenderPath.  config["rand"]: optional `TrainRandom` (seeded parity runs)."""
        from .train import TrainRandom, patch_pixels
        if self.train_path is None:
            self.train(self.training)
            if self.train_path is None:
                from .train import TrainableRenderPath
                self.train_path = TrainableRenderPath(self._state_dict, self.device, self.renderer.lib)
        path = self.train_path
        path.rand_noise_std = float(config.get("rand_noise_std", 0.0)) if self.training or "rand_noise_std" in config else 0.0
        rand = config.get("rand") or TrainRandom()
        width, height = int(cam_tar.get("width", cam_in["width"])), int(cam_tar.get("height", cam_in["height"]))
        if self.training and "msk" in config:
            out_h, out_w = self.train_out_h, self.train_out_w
            grids = patch_pixels(config["msk"][0].squeeze().bool().cpu(), width, height, out_h, out_w, rand)
        else:
            step = 2 ** (level - 1)
            sx, sy = (int(stride.reshape(-1, 2)[0, 0]), int(stride.reshape(-1, 2)[0, 1])) if isinstance(stride, torch.Tensor) else (int(stride),) * 2
            out_w, out_h = width // step, height // step
            ys, xs = torch.meshgrid(torch.arange(0, height, step), torch.arange(0, width, step), indexing="ij")
            grids = torch.stack([xs + sx, ys + sy], -1).reshape(-1, 2)
        frame = dict(img=img_in, cam_in=cam_in, targets=targets, sp_data=sp_data, feat_geo=feat_geo, feat_tex=feat_tex,
                     src_foreground_mask=config["src_foreground_mask"])
        path.set_frame(frame)
        o = path.render(cam_tar, config["bounds"], grids, rand=rand, training=self.training, uniform=bool(config.get("uniform", False)),
                        fine=bool(config.get("fine", False)), S_c=config.get("sample_per_ray_c", 64), S_f=config.get("sample_per_ray_f", 64),
                        znear=cam_tar.get("znear", cam_in["znear"]), zfar=cam_tar.get("zfar", cam_in["zfar"]))
        img = lambda t, c: t.reshape(out_h, out_w, c).permute(2, 0, 1)[None]
        out = {"tex_fg": img(o["tex_fg"], 3), "depth": o["depth"].reshape(1, out_h, out_w), "alpha": o["alpha"].reshape(1, out_h, out_w)}
        if "tex_fg_fine" in o:
            out.update({"tex_fg_fine": img(o["tex_fg_fine"], 3), "depth_fine": o["depth_fine"].reshape(1, out_h, out_w),
                        "alpha_fine": o["alpha_fine"].reshape(1, out_h, out_w), "sdf": o["sdf"].reshape(1, out_h, out_w)})
        index = (grids[:, 0] + grids[:, 1] * width).long().to(self.device)
        with torch.no_grad():
            if tar_img is not None:
                t = tar_img.reshape(*tar_img.shape[:2], -1).to(self.device)
                out["tar_img"] = torch.gather(t, 2, index[None, None].expand(t.shape[0], 3, -1)).view(t.shape[0], 3, out_h, out_w)
        out["vert_vis"] = path.vert_vis[:, :, None]
        return out

    # ---------------------------------------------------------------- render_pifu_nerf (src/model.py:1027-1100)
    @staticmethod
    def render_pifu_nerf(self, net, img_in, cam_in, hand_type, targets, cam_tar, level=5, sp_data={}, bkg_emb=None,
                         camcenter=None, objcenter=None, tar_img=None, **config):
        """Full-resolution image.  The reference renders stride^2 interleaved sub-images one after another and
        pixel-shuffles them together; every pixel is independent, so here all H*W rays go through one call."""
        n_views = img_in.shape[0]
        feat_geo = config.pop("feat_geo", None) or net.attach_geo_feat(img_in, return_val=True)
        feat_tex = config.pop("feat_tex", None)
        if feat_tex is None:
            feat_tex = net.attach_tex_feat(img_in, return_val=True)
        out = net.batch_render_pifu_nerf(net, img_in, cam_in, hand_type, targets, n_views, cam_tar, 1, 0, tar_img,
                                         feat_geo, feat_tex, None, sp_data, objcenter, **config)
        ret = {k: v[0] for k, v in out.items() if v is not None and k != "vert_vis" and v.dim() >= 3}
        ret = {k: (v[None] if v.dim() == 2 else v) for k, v in ret.items()}
        cam = cam_tar
        vert3d = targets["vert_world"].to(net.device)
        KRT = cam["KRT"].to(net.device)
        vimg = vert3d @ KRT[:, :3, :3].transpose(1, 2) + KRT[:, :3, 3][:, None]
        ret["vert_xy"] = vimg[..., :2] / (vimg[..., 2:3] + 1e-8)
        ret["vert_vis"] = out["vert_vis"]
        return ret
