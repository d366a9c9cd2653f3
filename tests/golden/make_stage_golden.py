"""Generates tests/golden/stages_v3.npz from the REFERENCE's own stage modules (build container only):

    python tests/golden/make_stage_golden.py

`GeoVisFusion`, `TexVisFusion`, `MLPUNetFusion`, `IBRRenderingHead`, `SpatialEncoder`, `feat_sample` and `KNN_vis` are
imported from /root/reference through oracle/ref_import.py (V-generalised flavour; the pytorch3d
knn_points entry point is served by oracle/geom.py), loaded with the seeded 'stress' weights (vanerf_b200.weights) and
called on the seeded inputs of tests/stage_inputs.py.  Only the outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import stage_inputs as SI
    from oracle import ref_import
    from vanerf_b200 import weights
    ns = ref_import.load(patched=True)
    sd = weights.init_state_dict(SI.H, SI.W, mode="stress")
    net = ref_import.build_net(ns, SI.H, SI.W, weights.to_torch(sd))
    d = {k: torch.from_numpy(v) for k, v in SI.make().items()}
    V = SI.V
    out = {}
    with torch.no_grad():
        out["feat_sample"] = ns.utils.feat_sample(d["g0"], d["uv"]).numpy()
        vf = ns.utils.feat_sample(d["g1"], d["vert_xy"])
        a, b, va, vb = ns.networks.KNN_vis(d["v"], d["vert"], vf, d["vert_vis"], 1)
        out["knn_a"], out["knn_b"], out["knn_va"], out["knn_vb"] = a.numpy(), b.numpy(), va.numpy(), vb.numpy()
        g = net.geo_vis_fusion(d["vert_xy"], [d["g0"], d["g1"]], [d["px64"][:, None], d["px8"][:, None]], d["vert"], d["v"], d["vert_vis"],
                               d["query_vis"], None, d["query_sdf"])
        out["geo64"], out["geo8"] = g[0].reshape(V, SI.N, 64).numpy(), g[1].reshape(V, SI.N, 8).numpy()
        t = net.tex_vis_fusion(d["vert_xy"], d["tex"], d["ft_xy"], d["vert"], d["v"], d["vert_vis"], d["query_vis"], d["img_xy"], d["img"], d["latent24"])
        out["tex40"] = t.numpy()
        o, valid, x_view, x_pool = net.mlp_geo(d["pe"], [d["f64"], d["f8"]], d["a"], d["w"])
        out["mlp_out"], out["mlp_valid"], out["mlp_view"], out["mlp_pool"] = o.numpy(), valid.numpy(), x_view.numpy(), x_pool.numpy()
        out["ibr_rgb"] = net.mlp_tex(d["rgb_feats"], d["ray_diffs"], d["proj_mask"]).numpy()
        enc = net.sp_encoder
        pts = d["v"][:1]
        out["sp"] = enc(KRT=d["extrin"], v=d["v"], pts=pts, n_view=V, z=None, xy=None, extrin=d["extrin"], kpt3d=d["kpt3d"]).numpy()
        out["sp_dim"] = np.int64(enc.get_dim())
        out["pe3"] = ns.spatial.SpatialEncoder.position_embedding(d["v"], 3).numpy()
    np.savez_compressed(os.path.join(HERE, "stages_v3.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
