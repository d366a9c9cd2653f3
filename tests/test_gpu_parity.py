"""Parity tests proper: the CUDA library through the C ABI on a B200 against the oracle (same seeded inputs), against
the golden vectors of the reference itself, and - at BASELINE.json's full size - through size-independent properties."""
import numpy as np
import pytest
import torch

import parity
from vanerf_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,V,mode,layout,npix", [
    (256, 256, 1, "ref", "narrow", 8),
    (256, 256, 1, "stress", "narrow", 8),
    (512, 334, 3, "ref", "narrow", 10),
    (512, 334, 3, "stress", "narrow", 12),
    (512, 334, 3, "stress", "bvv", 12),
    (512, 334, 4, "stress", "narrow", 6),
    (512, 334, 2, "stress", "bvv", 6),
])
def test_cuda_stages_match_oracle_fp32(cuda_lib, H, W, V, mode, layout, npix):
    sc, inp, sd = parity.build_case(H, W, V, mode=mode, layout=layout)
    r, vert_vis = parity.make_renderer(inp, sd, "cuda:0")
    pix = parity.lattice_pixels(H, W, npix)
    errs, oo, ot = parity.check_all(r, vert_vis, inp, sd, pix, precision=L.FP32)
    assert ot["valid"].any() and not ot["valid"].all()
    parity.check_render_rays(r, inp, oo, pix)
    assert r.launches > 0
    print("max-abs errors", {k: f"{v:.2e}" for k, v in errs.items()})


@pytest.mark.parametrize("name", ["v1_256_ref", "v3_512x334_ref", "v3_512x334_str", "v3_bvv_str"])
def test_cuda_surface_matches_reference_golden(cuda_lib, name):
    from test_model_surface import run_surface_case
    run_surface_case(name, "cuda:0", None)


def test_cuda_ragged_and_tiny_sizes(cuda_lib):
    from oracle import oracle_torch as OT
    sc, inp, sd = parity.build_case(256, 256, 2, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cuda:0")
    orc = OT.Oracle(sd, inp)
    for n_rays, S in [(1, 8), (7, 24), (33, 65), (130, 64)]:
        pix = parity.lattice_pixels(256, 256, 12)[:n_rays]
        ot = {}
        orc.render(fine=False, pixels=pix, S_c=S, taps=ot)
        tar = r.make_target(inp["cam_tar"], inp["bounds"])
        rays, z = r.sample_rays(tar, torch.from_numpy(pix), S)
        parity.assert_exact("z", z.cpu().numpy(), ot["z"])
        geo = r.geom_query(tar, rays, z)
        parity.assert_exact("face", geo["face"].cpu().numpy().astype(np.int64), ot["geo"]["face"])
        rgba, valid, raw = r.shade(tar, rays, z, geo)
        parity.assert_exact("valid", valid.cpu().numpy() > 0, ot["valid"])
        parity.assert_close("rgba", rgba.cpu().numpy(), ot["rgba"], parity.TOL_FP32)


def test_cuda_full_view_properties(cuda_lib):
    """Config B size (334x512, V=3, 64+64): properties that do not need the oracle at full size + an oracle spot check."""
    H, W, V = 512, 334, 3
    sc, inp, sd = parity.build_case(H, W, V, mode="stress")
    r, vert_vis = parity.make_renderer(inp, sd, "cuda:0")
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    pix = torch.stack([xs, ys], -1).reshape(-1, 2)
    R = pix.shape[0]
    assert R == 171008
    rays, z = r.sample_rays(tar, pix, 64)
    zc = z.cpu().numpy()
    assert np.isfinite(zc).all() and (np.diff(zc, axis=1) > 0).all(), "coarse depths strictly increasing"
    # linearity of the depth table: z = near + (far-near)*t reproduced on the host bit for bit
    rn = rays.cpu().numpy()
    t = torch.linspace(0, 1, 64).numpy()
    assert np.array_equal(zc, (rn[:, 3:4] + (rn[:, 4:5] - rn[:, 3:4]) * t[None]).astype(np.float32))
    oc, of = r.render_rays(tar, pix, 64, 64, True)
    oc, of = oc.cpu().numpy(), of.cpu().numpy()
    assert np.isfinite(oc).all() and np.isfinite(of).all()
    # the last interval is 1e10 (B-7), so alpha saturates at 1 wherever the last sample's density has not underflowed;
    # rays that stay ~0.5 m away from the hands end at 1 - exp(-sigmoid(-sdf/beta)/beta * 1e10) ~ 0.993 (oracle: same)
    for a in (oc[:, 4], of[:, 4]):
        assert a.max() < 1 + 1e-4 and a.min() > 0.95, "alpha in (0.95, 1]"
        assert (np.abs(a - 1) < 1e-4).mean() > 0.2, "alpha == 1 on the rays that come close to the hands"
    assert (of[:, 3] > 0.5).all() and (of[:, 3] < 1.6).all(), "depth inside the frustum"
    # the chunked full-view call must equal the same rays rendered as a small batch (chunk-boundary independence)
    sel = np.random.RandomState(0).choice(R, 96, replace=False)
    oc2, of2 = r.render_rays(tar, pix[sel], 64, 64, True)
    assert np.array_equal(oc2.cpu().numpy(), oc[sel]) and np.array_equal(of2.cpu().numpy(), of[sel])
    # oracle spot check on those rays
    from oracle import oracle_torch as OT
    oo = OT.Oracle(sd, inp).render(fine=True, pixels=pix[sel].numpy())
    parity.assert_close("full-view tex_fg vs oracle", oc[sel, :3], oo["tex_fg"], 1e-3)
    parity.assert_close("full-view tex_fg_fine vs oracle", of[sel, :3], oo["tex_fg_fine"], 5e-3)


def test_cuda_frame_switch_and_idempotence(cuda_lib):
    """Two frames through one context (render_dynamic usage): results depend only on the current frame."""
    sc0, inp0, sd = parity.build_case(256, 256, 3, mode="stress", frame=0)
    sc1, inp1, _ = parity.build_case(256, 256, 3, mode="stress", frame=5)
    r, _ = parity.make_renderer(inp0, sd, "cuda:0")
    pix = torch.from_numpy(parity.lattice_pixels(256, 256, 8))
    tar = r.make_target(inp0["cam_tar"], inp0["bounds"])
    a0 = r.render_rays(tar, pix)[1].clone()
    mv = lambda t: t.to("cuda:0")
    r.set_frame(mv(inp1["img"]), inp1["cam_in"], inp1["targets"], inp1["sp_data"], [mv(t) for t in inp1["feat_geo"]], mv(inp1["feat_tex"]), mv(inp1["src_foreground_mask"]))
    a1 = r.render_rays(r.make_target(inp1["cam_tar"], inp1["bounds"]), pix)[1].clone()
    r.set_frame(mv(inp0["img"]), inp0["cam_in"], inp0["targets"], inp0["sp_data"], [mv(t) for t in inp0["feat_geo"]], mv(inp0["feat_tex"]), mv(inp0["src_foreground_mask"]))
    a2 = r.render_rays(tar, pix)[1]
    assert torch.equal(a0, a2) and not torch.equal(a0, a1)


def test_cuda_errors_are_loud(cuda_lib):
    from vanerf_b200.renderer import Renderer
    r = Renderer("cuda:0")
    with pytest.raises(L.VanerfError):
        r.lib.check(r.ctx, r.lib.dll.vanerf_geom_query(r.ctx, None, None, None, 1, 1, None, None, None, None, None, None), "geom")


def test_cuda_render_sequence_frames_are_independent(cuda_lib):
    """render_dynamic driver (vanerf_b200.dynamic, BASELINE.json configs[3]): every frame of a sequence rendered through
    one network object equals that frame rendered alone, whatever came before it; the round-robin rank shards cover
    the sequence (the ranks of a 2-GPU job, run here one after the other on cuda:0)."""
    from vanerf_b200 import dynamic, synthetic, weights
    from vanerf_b200.model import VANeRF
    H, W, V, F = 256, 256, 3, 3
    sd = weights.init_state_dict(H, W, mode="stress")
    frames = {f: synthetic.to_torch(synthetic.make_scene(H, W, V, frame=f)) for f in range(F)}
    K = frames[0]["cam_tar"]["K"][0, :3, :3]
    cams = dynamic.orbit_cameras(F, K, width=W, height=H)
    cfg = dict(fine=True, sample_per_ray_c=16, sample_per_ray_f=16)

    def make_net():
        n = VANeRF(device=torch.device("cuda:0"), precision="fp32").eval()
        n.load_state_dict(sd)
        return n

    net = make_net()
    ids, seq = dynamic.render_sequence(net, lambda f: frames[f], F, lambda f: [cams[f]], 0, 1, **cfg)
    assert ids == list(range(F))
    for f in range(F):
        alone = dynamic.render_novel_views(make_net(), dynamic.to_device_frame(frames[f], "cuda:0"), [cams[f]], **cfg)
        assert torch.equal(seq[f], alone), f"frame {f} depends on its predecessors"
        assert torch.isfinite(seq[f]).all()
    assert not torch.equal(seq[0], seq[1])
    # the two shards of a 2-rank job
    got = {}
    for rank in range(2):
        ids_r, out_r = dynamic.render_sequence(make_net(), lambda f: frames[f], F, lambda f: [cams[f]], rank, 2, **cfg)
        got.update(dict(zip(ids_r, out_r)))
    assert sorted(got) == list(range(F)) and all(torch.equal(got[f], seq[f]) for f in range(F))


@pytest.mark.parametrize("precision", [L.FP32, L.BF16])
def test_cuda_coarse_reuse_is_bit_identical(cuda_lib, precision):
    """vanerf_set_reuse_coarse on the real library, both precision paths: same output bits with a third fewer network
    evaluations per ray (the merged fine set contains the coarse depths bit for bit; a sample's result does not depend
    on which tile or launch evaluates it)."""
    sc, inp, sd = parity.build_case(512, 334, 3, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cuda:0")
    pix = torch.from_numpy(parity.lattice_pixels(512, 334, 24))
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    oc0, of0 = r.render_rays(tar, pix, 64, 64, True, precision)
    oc0, of0 = oc0.clone(), of0.clone()
    n0 = r.launches
    r.set_reuse_coarse(True)
    oc1, of1 = r.render_rays(tar, pix, 64, 64, True, precision)
    r.set_reuse_coarse(False)
    assert r.launches > n0
    assert torch.equal(oc0, oc1) and torch.equal(of0, of1)
    assert torch.isfinite(of1).all() and float(of1[:, :3].abs().max()) > 0


@pytest.mark.parametrize("precision", [L.FP32, L.BF16])
def test_cuda_geometry_reuse_is_bit_identical(cuda_lib, precision):
    """vanerf_set_reuse_geometry (default on): the fine pass queries the mesh for the new depths only and keeps the coarse
    pass's sdf / nearest vertex / sample visibility for the coarse depths of the merged set.  Same output bits as querying
    the mesh for all merged samples; the networks evaluate every merged sample either way."""
    sc, inp, sd = parity.build_case(512, 334, 3, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cuda:0")
    pix = torch.from_numpy(np.random.RandomState(3).randint(0, [334, 512], size=(2500, 2)).astype(np.int32))
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    r.set_reuse_geometry(True)
    oc0, of0 = r.render_rays(tar, pix, 64, 64, True, precision)
    oc0, of0 = oc0.clone(), of0.clone()
    r.set_reuse_geometry(False)
    oc1, of1 = r.render_rays(tar, pix, 64, 64, True, precision)
    r.set_reuse_geometry(True)
    r.finish()
    assert torch.equal(oc0, oc1) and torch.equal(of0, of1)
    assert torch.isfinite(of1).all() and float(of1[:, :3].abs().max()) > 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cuda_partitioned_view_equals_single_gpu_image(cuda_lib, world):
    """vanerf_b200.dist.render_view: the image assembled from the `world` interleaved rank tiles (ranks run here one after the
    other on cuda:0, the all_gather replaced by a list) is bit-identical to the single-GPU image, in the reference's pixel order."""
    from vanerf_b200 import dist as D
    H, W, V = 64, 48, 3
    sc, inp, sd = parity.build_case(H, W, V, mode="stress")
    r, _ = parity.make_renderer(inp, sd, "cuda:0")
    tar = r.make_target(inp["cam_tar"], inp["bounds"])
    full = D.render_view(r, tar, H, W, 0, 1, 16, 16, True, L.FP32)
    assert full.shape == (H, W, 16) and torch.isfinite(full).all()
    tiles = {}
    for rank in range(world):
        D.render_view(r, tar, H, W, rank, world, 16, 16, True, L.FP32, gather=lambda t, rank=rank: tiles.__setitem__(rank, t.clone()) or [t] * world)
    img = D.assemble([tiles[k] for k in range(world)], H, W, world)
    assert torch.equal(img, full)
    # pixel order: row y, column x of the image is target pixel (x, y)
    ys, xs = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    oc, of = r.render_rays(tar, torch.stack([xs, ys], -1).reshape(-1, 2), 16, 16, True, L.FP32)
    assert torch.equal(full[..., :8].reshape(-1, 8), of) and torch.equal(full[..., 8:].reshape(-1, 8), oc)


def test_cuda_two_rank_nccl_image_equals_single_gpu_image(cuda_lib):
    """Same through two real ranks and NCCL (needs 2 GPUs; skipped on a 1-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(root, "tools", "dist_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "DIST_CHECK_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_cuda_frame_key_survives_recycled_addresses(cuda_lib):
    """frame 0's tensors are freed and frame 1 is allocated with the same shapes (the caching allocator hands out
    the same addresses); batch_render_pifu_nerf must render frame 1, not the cached frame 0."""
    from vanerf_b200 import synthetic, weights
    from vanerf_b200.model import VANeRF
    H, W, V = 256, 256, 3
    net = VANeRF(device="cuda:0", precision="fp32").eval()
    net.load_state_dict(weights.init_state_dict(H, W, mode="stress"))
    pix = torch.from_numpy(parity.lattice_pixels(H, W, 6))[None]

    def render(frame):
        inp = synthetic.to_torch(synthetic.make_scene(H, W, V, frame=frame), "cuda:0")
        out = VANeRF.batch_render_pifu_nerf(net, inp["img"], inp["cam_in"], inp["hand_type"], inp["targets"], V, inp["cam_tar"], 1, 0, None,
                                            inp["feat_geo"], inp["feat_tex"], None, dict(inp["sp_data"]), inp["objcenter"], fine=True, uniform=True,
                                            sample_per_ray_c=16, sample_per_ray_f=16, src_foreground_mask=inp["src_foreground_mask"],
                                            bounds=inp["bounds"], pixel_override=pix)
        ptr = inp["img"].data_ptr()
        return out["tex_fg_fine"].clone(), ptr

    a0, p0 = render(0)
    a1, p1 = render(3)                       # frame 0's tensors are gone: same shapes, (typically) same addresses
    fresh = VANeRF(device="cuda:0", precision="fp32").eval()
    fresh.load_state_dict(weights.init_state_dict(H, W, mode="stress"))
    inp = synthetic.to_torch(synthetic.make_scene(H, W, V, frame=3), "cuda:0")
    ref = VANeRF.batch_render_pifu_nerf(fresh, inp["img"], inp["cam_in"], inp["hand_type"], inp["targets"], V, inp["cam_tar"], 1, 0, None,
                                        inp["feat_geo"], inp["feat_tex"], None, dict(inp["sp_data"]), inp["objcenter"], fine=True, uniform=True,
                                        sample_per_ray_c=16, sample_per_ray_f=16, src_foreground_mask=inp["src_foreground_mask"],
                                        bounds=inp["bounds"], pixel_override=pix)["tex_fg_fine"]
    assert torch.equal(a1, ref) and not torch.equal(a0, a1)


def test_cuda_render_pifu_nerf_matches_oracle_and_reference_golden(cuda_lib):
    """VANeRF.render_pifu_nerf (src/model.py:1027-1100) on the CUDA library: dict layout + oracle check at 256 x 256, and the
    unpatched reference's own golden (V = 1, level 5 lattice) read back from the full image."""
    from test_model_surface import run_render_pifu_nerf_case, GOLD
    import os
    out = run_render_pifu_nerf_case("cuda:0", None, 64, 48, 3, 32, 32)
    assert torch.isfinite(out["tex_fg_fine"]).all()
    from vanerf_b200 import synthetic, weights
    from vanerf_b200.model import VANeRF
    g = np.load(os.path.join(GOLD, "v1_256_ref.npz"))
    H = W = 256
    inp = synthetic.to_torch(synthetic.make_scene(H, W, 1, layout=str(g["layout"])), "cuda:0")
    net = VANeRF(device="cuda:0", precision="fp32").eval()
    net.load_state_dict(weights.init_state_dict(H, W, mode=str(g["mode"])))
    net.attach_im_feat(feat_geo=inp["feat_geo"], feat_tex=inp["feat_tex"])
    cam_tar = dict(inp["cam_tar"])
    cam_tar["KRT"] = cam_tar["K"] @ cam_tar["RT"]
    out = VANeRF.render_pifu_nerf(None, net, inp["img"], inp["cam_in"], inp["hand_type"], inp["targets"], cam_tar, 5, dict(inp["sp_data"]),
                                  None, None, inp["objcenter"], None, fine=True, uniform=True, sample_per_ray_c=64, sample_per_ray_f=64,
                                  src_foreground_mask=inp["src_foreground_mask"], bounds=inp["bounds"])
    step = 2 ** (int(g["level"]) - 1)                      # the golden rays: the level-5 lattice, stride 0
    sub = lambda t: t[:, ::step, ::step].reshape(t.shape[0], -1).T.cpu().numpy()
    parity.assert_close("render_pifu_nerf tex_fg vs reference golden", sub(out["tex_fg"]), g["tex_fg"], 1e-3)
    parity.assert_close("render_pifu_nerf alpha vs reference golden", sub(out["alpha"])[:, 0], g["alpha"], 1e-3)
    parity.assert_close("render_pifu_nerf tex_fg_fine vs reference golden", sub(out["tex_fg_fine"]), g["tex_fg_fine"], parity.TOL_E2E_FINE_FP32)
