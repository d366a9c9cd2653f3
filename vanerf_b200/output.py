"""Output side of a render: image conversion, PNG files, the evaluator's metrics.

    tensor -> uint8 image, `*.pred.png` / `*.gt.png`       src/model.py:237-275 (save_test_image), :182-206 (render_dynamic frames)
    Evaluator.compute_score (mse, psnr, ssim, crop + files)   src/evaluator.py:14-47,84-114

Host-side glue, not a hot path: the clamp / scale / uint8 conversion runs on the device (one 0.5 MB read-back per view instead of
6 MB of floats), everything else is numpy.  Neither OpenCV, imageio nor scikit-image exist in this image, so the PNG encoder is the
20 lines below (zlib, filter 0) and SSIM restates the published algorithm `skimage.metrics.structural_similarity` implements with its
defaults (Wang et al. 2004: 7x7 uniform window, K1 = 0.01, K2 = 0.03, sample covariance, border of 3 cropped, mean over channels;
data range 2.0 for float images as in the scikit-image release the reference pins, `requirements.txt`).  PARITY UNPINNED against
scikit-image itself (not installable offline); `tests/test_output.py` checks it against a scipy.ndimage restatement of the same
formulas.  LPIPS needs the AlexNet + linear-head weights of the `lpips` package, which are not available offline: `lpips` is None.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Optional

import numpy as np
import torch


def to_uint8_image(img: torch.Tensor) -> np.ndarray:
    """(3,H,W) or (1,3,H,W) float in [0,1] (values outside are clamped, src/model.py:583) -> (H,W,3) uint8, RGB.
    The reference truncates (`(x * 255.).astype(np.uint8)`), so does this."""
    if img.dim() == 4:
        img = img[0]
    return (img.detach().clamp(0.0, 1.0) * 255.0).to(torch.uint8).permute(1, 2, 0).contiguous().cpu().numpy()


def write_png(path: str, img: np.ndarray) -> None:
    """(H,W,3) or (H,W) uint8 -> 8-bit PNG (RGB / grey), no external library."""
    a = np.ascontiguousarray(img, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    if c not in (1, 3):
        raise ValueError("write_png: 1 or 3 channels")
    raw = np.concatenate([np.zeros((h, 1), np.uint8), a.reshape(h, w * c)], axis=1).tobytes()        # filter type 0 per row

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2 if c == 3 else 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def read_png(path: str) -> np.ndarray:
    """Reads back what `write_png` wrote (8-bit, filter 0 only): used by the tests and by `Evaluator` round trips."""
    b = open(path, "rb").read()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w = 8, b"", 0
    while pos < len(b):
        n, tag = struct.unpack(">I", b[pos:pos + 4])[0], b[pos + 4:pos + 8]
        data = b[pos + 8:pos + 8 + n]
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", data[:10])
            c = 3 if ctype == 2 else 1
        elif tag == b"IDAT":
            idat += data
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * c)
    assert not rows[:, 0].any(), "read_png: only filter type 0"
    out = rows[:, 1:].reshape(h, w, c)
    return out if c == 3 else out[:, :, 0]


def bounding_rect(mask: np.ndarray):
    """cv2.boundingRect of the non-zero pixels: (x, y, w, h); an empty mask gives (0, 0, 0, 0)."""
    ys, xs = np.nonzero(mask)
    if ys.size == 0:
        return 0, 0, 0, 0
    return int(xs.min()), int(ys.min()), int(xs.max() - xs.min() + 1), int(ys.max() - ys.min() + 1)


def psnr(pred: np.ndarray, gt: np.ndarray) -> float:
    """src/evaluator.py:14-18: -10 log10(mean squared error), images in [0,1]."""
    mse = float(np.mean((np.asarray(pred, np.float64) - np.asarray(gt, np.float64)) ** 2))
    return float(-10.0 * np.log(mse) / np.log(10.0))


def ssim(pred, gt, data_range: float = 2.0, win: int = 7) -> float:
    """Mean structural similarity of two (H,W,C) float images, scikit-image defaults (see the module docstring)."""
    x = torch.as_tensor(np.asarray(pred), dtype=torch.float64).permute(2, 0, 1)[None]
    y = torch.as_tensor(np.asarray(gt), dtype=torch.float64).permute(2, 0, 1)[None]
    if min(x.shape[-2:]) < win:
        raise ValueError("ssim: image smaller than the 7x7 window")
    box = lambda t: torch.nn.functional.avg_pool2d(t, win, stride=1)        # valid region = skimage's filtered image minus its border
    npix = win * win
    cov = npix / (npix - 1.0)
    ux, uy = box(x), box(y)
    vx, vy, vxy = cov * (box(x * x) - ux * ux), cov * (box(y * y) - uy * uy), cov * (box(x * y) - ux * uy)
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux * ux + uy * uy + c1) * (vx + vy + c2))
    return float(s.mean())


class Evaluator:
    """src/evaluator.py: crops prediction and ground truth to the bounding rectangle of `mask_at_box`, writes
    `<result_dir>/<human>/{pred,gt,input}/frame*_view*.png` and returns the metrics."""

    def __init__(self, result_dir: Optional[str] = None):
        self.result_dir = result_dir

    def compute_score(self, rgb_pred, rgb_gt, input_imgs, mask_at_box, human_idx, frame_index, view_index, **_unused):
        hwc = lambda t: t.squeeze(0).permute(1, 2, 0).detach().float().cpu().numpy()
        pred, gt = hwc(rgb_pred), hwc(rgb_gt)
        mask = mask_at_box.squeeze().detach().cpu().numpy()
        x, y, w, h = bounding_rect(mask)
        crop_p, crop_g = pred[y:y + h, x:x + w], gt[y:y + h, x:x + w]
        if self.result_dir is not None:
            base = os.path.join(self.result_dir, str(human_idx))
            for d in ("pred", "gt", "input"):
                os.makedirs(os.path.join(base, d), exist_ok=True)
            u8 = lambda a: (np.clip(a, 0.0, 1.0) * 255.0).astype(np.uint8)
            write_png(os.path.join(base, "gt", f"frame{frame_index}_view{view_index}_gt.png"), u8(crop_g))
            write_png(os.path.join(base, "pred", f"frame{frame_index}_view{view_index}.png"), u8(crop_p))
            ins = input_imgs.permute(0, 2, 3, 1).detach().float().cpu().numpy()
            for v in range(ins.shape[0]):          # the reference writes every source view to the SAME file name (src/evaluator.py:41-42): the last one stays
                write_png(os.path.join(base, "input", f"frame{frame_index}_t_0_view_{view_index}.png"), u8(ins[v][y:y + h, x:x + w]))
        return {"mse": float(np.mean((pred - gt) ** 2)), "psnr": psnr(pred, gt), "ssim": ssim(crop_p, crop_g), "lpips": None}


def save_test_image(dst_dir: str, tar_cam_id, rendered_img: Optional[torch.Tensor] = None, gt_img: Optional[torch.Tensor] = None,
                    mask: Optional[torch.Tensor] = None) -> None:
    """src/model.py:237-275: `<tar_cam_id>.pred.png`, `.gt.png`, `.mask.png` under dst_dir."""
    os.makedirs(dst_dir, exist_ok=True)
    if rendered_img is not None:
        write_png(os.path.join(dst_dir, f"{tar_cam_id}.pred.png"), to_uint8_image(rendered_img))
    if gt_img is not None:
        write_png(os.path.join(dst_dir, f"{tar_cam_id}.gt.png"), to_uint8_image(gt_img))
    if mask is not None:
        m = (mask.detach().float().squeeze().clamp(0, 1) * 255.0).to(torch.uint8).cpu().numpy()
        write_png(os.path.join(dst_dir, f"{tar_cam_id}.mask.png"), np.repeat(m[:, :, None], 3, axis=2))
