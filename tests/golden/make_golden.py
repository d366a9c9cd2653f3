"""Generates tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

    python tests/golden/make_golden.py            # writes all cases
    python tests/golden/make_golden.py v1_256_ref # one case (one reference flavour per process)

The reference modules are imported from /root/reference through oracle/ref_import.py (third-party stubs; kaolin /
pytorch3d entry points served by oracle/geom.py), the seeded synthetic scene (vanerf_b200.synthetic) and the
seeded weights (vanerf_b200.weights) are loaded into the reference `VANeRF`, and
`VANeRF.batch_render_pifu_nerf` (src/model.py:1103) is run with `fine=True, uniform=True`.  Inputs are NOT
stored (they are regenerated from the seeds); outputs and per-stage taps are.

Cases
  v1_256_ref      unpatched reference, V=1, 256x256, level 5 (256 rays), reference-like init
  v3_512x334_ref  V-generalised reference, V=3, 512x334, 12x12 target pixels, ref-like init
  v3_512x334_str  same with the 'stress' weight set (O(0.1-1) activations)
  v3_bvv_str      big-view-variation camera layout, stress weights
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = {
    "v1_256_ref": dict(V=1, H=256, W=256, level=5, patched=False, mode="ref", layout="narrow", npix=0),
    "v3_512x334_ref": dict(V=3, H=512, W=334, level=1, patched=True, mode="ref", layout="narrow", npix=12),
    "v3_512x334_str": dict(V=3, H=512, W=334, level=1, patched=True, mode="stress", layout="narrow", npix=12),
    "v3_bvv_str": dict(V=3, H=512, W=334, level=1, patched=True, mode="stress", layout="bvv", npix=12),
}


def case_pixels(npix):
    """Config A style pixel lattice (x = 5 + 10 i', y = 8 + 16 j' on a 32x32 grid), subsampled to npix x npix."""
    ii, jj = np.meshgrid(np.arange(npix), np.arange(npix), indexing="ij")
    return np.stack([(5 + 10 * (ii * 32 // npix)).ravel(), (8 + 16 * (jj * 32 // npix)).ravel()], 1).astype(np.int64)


def run_case(name):
    import torch
    from oracle import ref_import
    from vanerf_b200 import synthetic, weights
    c = CASES[name]
    ns = ref_import.load(patched=c["patched"])
    M = ns.model
    V, H, W = c["V"], c["H"], c["W"]
    sc = synthetic.make_scene(H, W, V, layout=c["layout"])
    inp = synthetic.to_torch(sc)
    sdnp = weights.init_state_dict(H, W, mode=c["mode"])
    net = ref_import.build_net(ns, H, W, weights.to_torch(sdnp))

    calls, r2o, bb = [], [], []
    oq, orr, ob = M.VANeRF.query, M.VANeRF.rgba2out, M.VANeRF.ray_bbox_intersection

    def q(self, pts, *a, **k):
        out = oq(self, pts, *a, **k)
        calls.append(dict(pts=pts.clone(), out=out[0].clone(), valid=out[1].clone(), query_vis=k['query_vis'].clone(),
                          query_sdf=k['query_sdf'].clone(), vert_vis=k['vert_vis'].clone(),
                          closest_face=k['closest_face'].clone()))
        return out

    def r(self_, rgba, z, sdf):
        o = orr(self_, rgba, z, sdf)
        r2o.append(dict(rgba=rgba.clone(), z=z.clone(), out=o))
        return o

    def b(bounds, orig, direct, **k):
        o = ob(bounds, orig, direct, **k)
        bb.append(dict(orig=orig.clone(), direct=direct.clone(), out=o))
        return o
    M.VANeRF.query, M.VANeRF.rgba2out, M.VANeRF.ray_bbox_intersection = q, staticmethod(r), staticmethod(b)
    extra = {}
    pix = None
    if c["npix"]:
        pix = case_pixels(c["npix"])
        extra["pixel_override"] = torch.from_numpy(pix)[None]
    with torch.no_grad():
        out = M.VANeRF.batch_render_pifu_nerf(
            net, inp['img'], inp['cam_in'], inp['hand_type'], inp['targets'], V, inp['cam_tar'], c["level"],
            torch.zeros(1, 2), None, inp['feat_geo'], inp['feat_tex'], None, dict(inp['sp_data']), inp['objcenter'],
            fine=True, uniform=True, sample_per_ray_c=64, sample_per_ray_f=64,
            src_foreground_mask=inp['src_foreground_mask'], bounds=inp['bounds'], **extra)
    f = lambda t: t.detach().cpu().numpy()
    S = 64
    R = bb[0]['direct'].shape[1]
    g = dict(
        case=np.array(name), V=V, H=H, W=W, level=c["level"], mode=np.array(c["mode"]), layout=np.array(c["layout"]),
        pixels=pix if pix is not None else np.zeros((0, 2), np.int64),
        cam_pos=f(bb[0]['orig'])[0, 0], cam_rays=f(bb[0]['direct'])[0],
        box_near=f(bb[0]['out'][0])[0, :, 0], box_far=f(bb[0]['out'][1])[0, :, 0], hit=f(bb[0]['out'][2])[0, :, 0],
        z=f(r2o[0]['z'])[0], sdf_mesh=f(calls[0]['query_sdf'])[0], query_vis=f(calls[0]['query_vis'])[:, :, 0],
        vert_vis=f(calls[0]['vert_vis'])[:, :, 0], closest_face=f(calls[0]['closest_face'])[0].astype(np.int32),
        valid=f(calls[0]['valid'])[0, :, 0], query_out=f(calls[0]['out'])[0], rgba=f(r2o[0]['rgba'])[0],
        contrib=f(r2o[0]['out'][3])[0],
        tex_fg=f(out['tex_fg'])[0].reshape(3, -1).T, depth=f(out['depth'])[0].reshape(-1), alpha=f(out['alpha'])[0].reshape(-1),
        z_fine=f(r2o[1]['z'])[0], sdf_mesh_fine=f(calls[1]['query_sdf'])[0], query_vis_fine=f(calls[1]['query_vis'])[:, :, 0],
        valid_fine=f(calls[1]['valid'])[0, :, 0], query_out_fine=f(calls[1]['out'])[0],
        tex_fg_fine=f(out['tex_fg_fine'])[0].reshape(3, -1).T, depth_fine=f(out['depth_fine'])[0].reshape(-1),
        alpha_fine=f(out['alpha_fine'])[0].reshape(-1), sdf=f(out['sdf'])[0].reshape(-1),
    )
    assert g['z'].shape == (R, S)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
    print(name, "written:", {k: (v.shape if hasattr(v, 'shape') else v) for k, v in g.items() if k in ('z', 'query_out', 'tex_fg_fine')})


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    if len(names) == 1:
        run_case(names[0])
    else:
        for n in names:       # one reference flavour (patched / unpatched) per process
            subprocess.check_call([sys.executable, os.path.abspath(__file__), n])
