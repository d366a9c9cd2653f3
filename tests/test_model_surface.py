"""The reference-facing call surface (vanerf_b200.model.VANeRF) against the golden vectors of the REFERENCE ITSELF.
CPU run = host-emulation build of the kernels (logic); the GPU run of the same checks is in test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

import parity
from vanerf_b200 import synthetic, weights
from vanerf_b200.model import VANeRF

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def run_surface_case(name, device, lib, precision="fp32", tol=1e-3, tol_fine=5e-3):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    V, H, W, level = int(g["V"]), int(g["H"]), int(g["W"]), int(g["level"])
    sc = synthetic.make_scene(H, W, V, layout=str(g["layout"]))
    inp = synthetic.to_torch(sc, device)
    net = VANeRF(device=device, precision=precision, lib=lib).eval()
    net.load_state_dict(weights.init_state_dict(H, W, mode=str(g["mode"])))
    extra = {}
    if g["pixels"].shape[0]:
        extra["pixel_override"] = torch.from_numpy(g["pixels"])[None]
    out = VANeRF.batch_render_pifu_nerf(
        net, inp["img"], inp["cam_in"], inp["hand_type"], inp["targets"], V, inp["cam_tar"], level, torch.zeros(1, 2), None,
        inp["feat_geo"], inp["feat_tex"], None, dict(inp["sp_data"]), inp["objcenter"], fine=True, uniform=True,
        sample_per_ray_c=64, sample_per_ray_f=64, src_foreground_mask=inp["src_foreground_mask"], bounds=inp["bounds"], **extra)
    f = lambda t: t.detach().cpu().numpy()
    parity.assert_exact("vert_vis vs reference", f(out["vert_vis"])[:, :, 0], g["vert_vis"])
    # per-pixel outputs: the north-star bar, literal (no scaling by the output range); fine pass: end-to-end bar (parity.py (3))
    parity.assert_close("tex_fg vs reference", f(out["tex_fg"])[0].reshape(3, -1).T, g["tex_fg"], tol)
    parity.assert_close("depth vs reference", f(out["depth"]).reshape(-1), g["depth"], tol)
    parity.assert_close("alpha vs reference", f(out["alpha"]).reshape(-1), g["alpha"], tol)
    parity.assert_close("tex_fg_fine vs reference", f(out["tex_fg_fine"])[0].reshape(3, -1).T, g["tex_fg_fine"], tol_fine)
    parity.assert_close("depth_fine vs reference", f(out["depth_fine"]).reshape(-1), g["depth_fine"], tol)
    assert out["tex_fg"].shape[1] == 3 and out["input_mask"].shape[1] == 1
    # VANeRF.query on the reference's own coarse sample positions (reconstructed from the golden rays / depths)
    R, S = g["z"].shape
    pts = (g["cam_pos"][None, None] + g["cam_rays"][:, None] * g["z"][:, :, None]).reshape(1, -1, 3).astype(np.float32)
    view = np.repeat(g["cam_rays"], S, 0)[None].astype(np.float32)
    qo, valid = net.query(torch.from_numpy(pts).to(device), inp["cam_in"], inp["hand_type"], inp["targets"], inp["feat_geo"],
                          inp["feat_tex"], n_views=V, sp_data=dict(inp["sp_data"]), tx_data={"img": inp["img"]},
                          view=torch.from_numpy(view).to(device), n_pts_samples=S,
                          src_foreground_mask=inp["src_foreground_mask"], bounds=inp["bounds"])
    parity.assert_exact("query valid vs reference", f(valid)[0, :, 0], g["valid"].astype(bool))
    qs = max(1.0, float(np.abs(g["query_out"]).max()))          # per-sample tap: relative to the output range (parity.py (2))
    parity.assert_close("query out vs reference", f(qo)[0], g["query_out"], tol * qs)
    return out


@pytest.mark.parametrize("name", ["v1_256_ref", "v3_512x334_str"])
def test_surface_matches_reference_golden_emulated(emul_lib, name):
    run_surface_case(name, "cpu", emul_lib)


def test_helpers_match_reference_semantics_emulated(emul_lib):
    """rgba2out / importance_sample with the reference's argument conventions."""
    from oracle import oracle_torch as OT
    rng = np.random.RandomState(1)
    net = VANeRF(device="cpu", lib=emul_lib)
    net.load_state_dict(weights.init_state_dict(256, 256, mode="stress"))
    R, S = 37, 64
    contrib = rng.uniform(0, 1, (R, S)).astype(np.float32) ** 3
    contrib /= contrib.sum(1, keepdims=True)
    z = np.sort(rng.uniform(0.7, 1.4, (R, S)).astype(np.float32), 1)
    z_mid = (np.float32(0.5) * (z[:, 1:] + z[:, :-1])).astype(np.float32)
    want = OT.Oracle.importance_sample(contrib[:, 1:-1], z_mid, 64)
    got = net.importance_sample(torch.from_numpy(contrib[:, 1:-1].copy())[None], torch.from_numpy(z_mid)[None], 64, uniform=True)
    parity.assert_exact("importance_sample (reference convention)", got[0].numpy(), want)
    near, far, hit = VANeRF.ray_bbox_intersection(torch.tensor([[[-.1, -.1, -.1], [.1, .1, .1]]]), torch.tensor([[[0., 0., -1.]]]),
                                                  torch.nn.functional.normalize(torch.tensor([[[0., 0., 1.], [0.05, 0., 1.], [1., 0., 0.]]]), dim=-1))
    assert hit[0, :, 0].tolist() == [True, True, False] and abs(float(near[0, 0, 0]) - 0.89) < 1e-5


def run_render_pifu_nerf_case(device, lib, H, W, V, S_c, S_f, precision="fp32"):
    """VANeRF.render_pifu_nerf (src/model.py:1027-1100): the reference's dict of (C,H,W) images, checked against the oracle."""
    from oracle import oracle_torch as OT
    sc = synthetic.make_scene(H, W, V)
    inp = synthetic.to_torch(sc, device)
    sd = weights.init_state_dict(H, W, mode="stress")
    net = VANeRF(device=device, precision=precision, lib=lib).eval()
    net.load_state_dict(sd)
    net.attach_im_feat(feat_geo=inp["feat_geo"], feat_tex=inp["feat_tex"])
    cam_tar = dict(inp["cam_tar"])
    cam_tar["KRT"] = cam_tar["K"] @ cam_tar["RT"]
    out = VANeRF.render_pifu_nerf(None, net, inp["img"], inp["cam_in"], inp["hand_type"], inp["targets"], cam_tar, 5, dict(inp["sp_data"]),
                                  None, None, inp["objcenter"], None, fine=True, uniform=True, sample_per_ray_c=S_c, sample_per_ray_f=S_f,
                                  src_foreground_mask=inp["src_foreground_mask"], bounds=inp["bounds"])
    shapes = {"tex_fg": (3, H, W), "depth": (1, H, W), "alpha": (1, H, W), "tex_fg_fine": (3, H, W), "depth_fine": (1, H, W),
              "alpha_fine": (1, H, W), "sdf": (1, H, W), "input_mask": (1, H, W), "img_in": (3, H, W)}
    for k, shp in shapes.items():
        assert tuple(out[k].shape) == shp, f"{k}: {tuple(out[k].shape)} vs {shp}"
    assert tuple(out["vert_xy"].shape) == (1, 1558, 2) and tuple(out["vert_vis"].shape) == (V, 1558, 1)
    f = lambda t: t.detach().cpu().numpy()
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    pix = np.stack([xs, ys], -1).reshape(-1, 2).astype(np.int64)
    sel = np.random.RandomState(5).choice(pix.shape[0], min(64, pix.shape[0]), replace=False)
    oo = OT.Oracle(sd, synthetic.to_torch(sc)).render(fine=True, pixels=pix[sel], S_c=S_c, S_f=S_f)
    img = lambda k: f(out[k]).reshape(out[k].shape[0], -1).T[sel]
    parity.assert_close("render_pifu_nerf tex_fg", img("tex_fg"), oo["tex_fg"], 1e-3)
    parity.assert_close("render_pifu_nerf alpha", img("alpha")[:, 0], oo["alpha"], 1e-3)
    parity.assert_close("render_pifu_nerf depth", img("depth")[:, 0], oo["depth"], 1e-3)
    parity.assert_close("render_pifu_nerf tex_fg_fine", img("tex_fg_fine"), oo["tex_fg_fine"], parity.TOL_E2E_FINE_FP32)
    assert np.array_equal(f(out["img_in"]), f(inp["img"][0])), "img_in = source image 0 gathered at the ray pixels"
    # vert_xy = target-camera projection of the mesh (src/model.py:1091-1097)
    v = inp["targets"]["vert_world"].cpu()
    KRT = cam_tar["KRT"].cpu()
    vimg = v @ KRT[:, :3, :3].transpose(1, 2) + KRT[:, :3, 3][:, None]
    assert torch.allclose(out["vert_xy"].cpu(), vimg[..., :2] / (vimg[..., 2:3] + 1e-8), atol=1e-4)
    return out


def test_render_pifu_nerf_emulated(emul_lib):
    # V = 3: with two source cameras placed symmetrically about the target (synthetic 'narrow' layout at V = 2) the head's
    # blending weights (e - min_v e) / (sum + 1e-8) are a 0/0 on the symmetry plane and amplify last-ulp differences of the ray
    # dot products into O(0.1) colour differences between ANY two fp32 implementations, the reference's own included.
    run_render_pifu_nerf_case("cpu", emul_lib, 24, 16, 3, 12, 12)


def test_partitioned_view_equals_single_image_emulated(emul_lib):
    """vanerf_b200.dist.render_view (partition -> render -> gather -> assemble) on the emulated kernels: the 2- and 4-rank
    images equal the single-rank image bit for bit."""
    from vanerf_b200 import _lib as L
    from vanerf_b200 import dist as D
    H, W, V = 16, 12, 2
    sc, inp, sd = parity.build_case(H, W, V, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cpu", emul_lib)
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    full = D.render_view(r, tar, H, W, 0, 1, 8, 8, True, L.FP32)
    for world in (2, 4):
        tiles = {}
        for rank in range(world):
            D.render_view(r, tar, H, W, rank, world, 8, 8, True, L.FP32, gather=lambda t, rank=rank: tiles.__setitem__(rank, t.clone()) or [t] * world)
        assert torch.equal(D.assemble([tiles[k] for k in range(world)], H, W, world), full)


def test_frame_key_sees_new_tensors_and_inplace_updates_emulated(emul_lib):
    """_ensure_frame: a new tensor object or an in-place update of ANY per-frame input triggers the per-frame setup; the very
    same objects do not."""
    H, W, V = 16, 12, 2
    sc = synthetic.make_scene(H, W, V)
    inp = synthetic.to_torch(sc)
    net = VANeRF(device="cpu", lib=emul_lib).eval()
    net.load_state_dict(weights.init_state_dict(H, W, mode="stress"))
    calls = []
    orig = net.renderer.set_frame
    net.renderer.set_frame = lambda *a, **k: calls.append(1) or orig(*a, **k)
    args = lambda d: (d["img"], d["cam_in"], d["targets"], d["sp_data"], d["feat_geo"], d["feat_tex"], d["src_foreground_mask"])
    net._ensure_frame(*args(inp))
    net._ensure_frame(*args(inp))
    assert len(calls) == 1
    inp["feat_geo"][0].mul_(1.0)                       # in-place update of a feature map (not img / vert_world)
    net._ensure_frame(*args(inp))
    assert len(calls) == 2
    inp2 = synthetic.to_torch(sc)                      # equal values, new tensor objects
    net._ensure_frame(*args(inp2))
    assert len(calls) == 3
