"""render_dynamic: a sequence of frames, each with its own mesh / source images / feature maps, rendered from a camera
path (BASELINE.json configs[3]; reference: render_dynamic.py:24-33 -> VANeRFLightningModule.render_video
src/model.py:141-207 -> render_novel_views :514-545, cameras from get_360cameras src/utils.py:63-134).

The reference walks frames x cameras sequentially on one GPU.  Frames are independent (per-frame state = mesh BVH, vertex
visibility, vertex tables, bf16 maps), so here they are dealt round-robin to the ranks: rank r renders frames
r, r + G, r + 2G, ... including their per-frame setup, with no collective inside the path; the finished images are
gathered once at the end (`gather_frames`).  One process per GPU.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from .model import VANeRF


def frames_for_rank(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin frame partition: frame f belongs to rank f mod world."""
    return list(range(rank, n_frames, world))


def orbit_cameras(n: int, K: torch.Tensor, radius: float = 1.0, elevation_deg: float = 0.0, width: int = 334, height: int = 512,
                  znear: float = 0.71, zfar: float = 1.42) -> List[Dict]:
    """n target cameras on a 360-degree path around the origin looking at it (the role of get_360cameras,
    src/utils.py:63-134), as the cam_tar dictionaries render_novel_views builds (src/model.py:521-528)."""
    from .synthetic import orbit_cam
    cams = []
    for i in range(n):
        Rt = torch.from_numpy(np.asarray(orbit_cam(360.0 * i / n, elevation_deg, radius), np.float32))
        RT = torch.eye(4)[None].clone()
        RT[0, :3, :4] = Rt
        K4 = torch.eye(4)[None].clone()
        K4[0, :3, :3] = K
        cams.append({"K": K4, "RT": RT, "KRT": torch.bmm(K4, RT), "width": width, "height": height, "nml_scale": 100.0,
                     "znear": znear, "zfar": zfar})
    return cams


def render_novel_views(net: VANeRF, frame: Dict, cameras: Sequence[Dict], **config) -> torch.Tensor:
    """All `cameras` of one frame -> (n_cam, 8, H, W) device tensor [rgb_fine 3 | depth_fine | alpha_fine | sdf | rgb_coarse r,g].
    `frame` = the reference-layout dictionaries of one time step (synthetic.to_torch / decode_batch): img, cam_in,
    targets, sp_data, feat_geo, feat_tex, src_foreground_mask, bounds, hand_type, objcenter (device tensors for the
    maps).  The per-frame setup runs once (first camera) and is reused by the others (VANeRF._ensure_frame)."""
    V = frame["img"].shape[0]
    outs = []
    for cam in cameras:
        o = VANeRF.batch_render_pifu_nerf(net, frame["img"], frame["cam_in"], frame["hand_type"], frame["targets"], V, cam, 1, 0, None,
                                          frame["feat_geo"], frame["feat_tex"], None, frame["sp_data"], frame.get("objcenter"),
                                          fine=config.get("fine", True), uniform=True,
                                          sample_per_ray_c=config.get("sample_per_ray_c", 64),
                                          sample_per_ray_f=config.get("sample_per_ray_f", 64),
                                          src_foreground_mask=frame["src_foreground_mask"], bounds=frame["bounds"])
        if "tex_fg_fine" in o:
            img = torch.cat([o["tex_fg_fine"][0], o["depth_fine"], o["alpha_fine"], o["sdf"], o["tex_fg"][0][:2]], 0)
        else:
            z = torch.zeros_like(o["depth"])
            img = torch.cat([o["tex_fg"][0], o["depth"], o["alpha"], z, o["tex_fg"][0][:2]], 0)
        outs.append(img)
    return torch.stack(outs)


def to_device_frame(frame_host: Dict, device) -> Dict:
    """Host (pinned) frame dictionary -> device copies of the tensors the path reads (async copies on the current stream)."""
    f = dict(frame_host)
    f["img"] = frame_host["img"].to(device, non_blocking=True)
    f["feat_tex"] = frame_host["feat_tex"].to(device, non_blocking=True)
    f["feat_geo"] = [t.to(device, non_blocking=True) for t in frame_host["feat_geo"]]
    m = frame_host["src_foreground_mask"]
    f["src_foreground_mask"] = (m.to(device, non_blocking=True) if m.dtype == torch.bool else m.to(device, non_blocking=True).bool())
    return f


def h2d_bytes(frame_host: Dict) -> int:
    ts = [frame_host["img"], frame_host["feat_tex"], frame_host["src_foreground_mask"]] + list(frame_host["feat_geo"])
    return int(sum(t.numel() * t.element_size() for t in ts))


def render_sequence(net: VANeRF, get_frame: Callable[[int], Dict], n_frames: int, cameras_of: Callable[[int], Sequence[Dict]],
                    rank: int = 0, world: int = 1, out_host: Optional[torch.Tensor] = None, **config):
    """Renders this rank's frames (round robin).  get_frame(f) returns the HOST dictionaries of frame f; each frame is
    copied to the device, set up and rendered from cameras_of(f).  Returns (frame ids, list of (n_cam, 8, H, W) device
    tensors); with `out_host` (pinned, (n_local, n_cam, 8, H, W)) the images are also copied back asynchronously."""
    ids = frames_for_rank(n_frames, rank, world)
    outs = []
    for k, f in enumerate(ids):
        fr = to_device_frame(get_frame(f), net.device)
        net.invalidate_frame()         # a new time step: never reuse the previous frame's setup
        img = render_novel_views(net, fr, cameras_of(f), **config)
        if out_host is not None:
            out_host[k].copy_(img, non_blocking=True)
        outs.append(img)
    return ids, outs


def gather_frames(local: Sequence[torch.Tensor], n_frames: int, rank: int, world: int):
    """All ranks' frame images -> list in frame order on every rank (one all_gather of equal-sized stacks; ranks with
    one frame fewer pad with zeros).  world == 1: identity."""
    if world == 1:
        return list(local)
    import torch.distributed as dist
    n_max = math.ceil(n_frames / world)
    shape = local[0].shape if len(local) else None
    if shape is None:
        raise ValueError("gather_frames: every rank needs at least one frame (n_frames >= world)")
    stack = torch.zeros((n_max,) + tuple(shape), dtype=local[0].dtype, device=local[0].device)
    for k, t in enumerate(local):
        stack[k] = t
    parts = [torch.empty_like(stack) for _ in range(world)]
    dist.all_gather(parts, stack)
    return [parts[f % world][f // world] for f in range(n_frames)]
