"""Output side: uint8 conversion, PNG round trip, PSNR / SSIM against restatements of the published formulas
(src/evaluator.py:14-47; scikit-image's structural_similarity defaults - the library itself is not installable offline)."""
import os

import numpy as np
import torch
from scipy import ndimage

from vanerf_b200 import output as O


def _ssim_scipy(x, y, data_range=2.0, win=7):
    """structural_similarity as scikit-image computes it: uniform_filter over the whole image, statistics with the sample
    covariance, border of (win - 1) / 2 cropped, mean over pixels then over channels."""
    vals = []
    for c in range(x.shape[2]):
        a, b = x[..., c].astype(np.float64), y[..., c].astype(np.float64)
        f = lambda t: ndimage.uniform_filter(t, size=win)
        npix = win * win
        cov = npix / (npix - 1.0)
        ux, uy = f(a), f(b)
        vx, vy, vxy = cov * (f(a * a) - ux * ux), cov * (f(b * b) - uy * uy), cov * (f(a * b) - ux * uy)
        c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
        s = ((2 * ux * uy + c1) * (2 * vxy + c2)) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2))
        p = (win - 1) // 2
        vals.append(s[p:-p, p:-p].mean())
    return float(np.mean(vals))


def test_png_round_trip_and_uint8_conversion(tmp_path):
    g = torch.Generator().manual_seed(0)
    img = torch.rand(3, 37, 53, generator=g) * 1.4 - 0.2                   # values outside [0,1] are clamped
    u8 = O.to_uint8_image(img)
    ref = (np.clip(img.permute(1, 2, 0).numpy(), 0.0, 1.0) * 255.0).astype(np.uint8)
    assert u8.dtype == np.uint8 and np.array_equal(u8, ref)
    p = os.path.join(tmp_path, "a.png")
    O.write_png(p, u8)
    assert np.array_equal(O.read_png(p), u8)
    O.write_png(p, u8[:, :, 0])
    assert np.array_equal(O.read_png(p), u8[:, :, 0])


def test_metrics_match_the_published_formulas():
    rs = np.random.RandomState(1)
    gt = rs.rand(64, 48, 3).astype(np.float32)
    pred = np.clip(gt + 0.05 * rs.randn(64, 48, 3).astype(np.float32), 0, 1)
    mse = np.mean((pred.astype(np.float64) - gt) ** 2)
    assert abs(O.psnr(pred, gt) - (-10.0 * np.log10(mse))) < 1e-9
    assert abs(O.ssim(pred, gt) - _ssim_scipy(pred, gt)) < 1e-9
    assert abs(O.ssim(gt, gt) - 1.0) < 1e-12
    assert O.bounding_rect(np.pad(np.ones((3, 5)), ((2, 4), (7, 1)))) == (7, 2, 5, 3) and O.bounding_rect(np.zeros((4, 4))) == (0, 0, 0, 0)


def test_evaluator_writes_the_references_files(tmp_path):
    rs = np.random.RandomState(2)
    gt = torch.from_numpy(rs.rand(1, 3, 40, 32).astype(np.float32))
    pred = (gt + 0.02 * torch.from_numpy(rs.randn(1, 3, 40, 32).astype(np.float32))).clamp(0, 1)
    mask = torch.zeros(40, 32)
    mask[5:30, 4:28] = 1
    ev = O.Evaluator(str(tmp_path))
    s = ev.compute_score(pred, gt, torch.rand(3, 3, 40, 32), mask, "cap0", 7, 2)
    assert set(s) == {"mse", "psnr", "ssim", "lpips"} and s["lpips"] is None and 0.5 < s["ssim"] <= 1.0 and s["psnr"] > 25
    base = os.path.join(tmp_path, "cap0")
    assert O.read_png(os.path.join(base, "pred", "frame7_view2.png")).shape == (25, 24, 3)
    assert O.read_png(os.path.join(base, "gt", "frame7_view2_gt.png")).shape == (25, 24, 3)
    assert os.path.exists(os.path.join(base, "input", "frame7_t_0_view_2.png"))
    O.save_test_image(os.path.join(tmp_path, "t"), "cam3", pred, gt, mask)
    assert sorted(os.listdir(os.path.join(tmp_path, "t"))) == ["cam3.gt.png", "cam3.mask.png", "cam3.pred.png"]
