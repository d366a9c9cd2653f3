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
    scale = max(1.0, float(np.abs(g["tex_fg"]).max()))
    parity.assert_close("tex_fg vs reference", f(out["tex_fg"])[0].reshape(3, -1).T, g["tex_fg"], tol * scale)
    parity.assert_close("depth vs reference", f(out["depth"]).reshape(-1), g["depth"], tol)
    parity.assert_close("alpha vs reference", f(out["alpha"]).reshape(-1), g["alpha"], tol)
    parity.assert_close("tex_fg_fine vs reference", f(out["tex_fg_fine"])[0].reshape(3, -1).T, g["tex_fg_fine"], tol_fine * scale)
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
    qs = max(1.0, float(np.abs(g["query_out"]).max()))
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
