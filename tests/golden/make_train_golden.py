"""Generates tests/golden/train_v3.npz from the REFERENCE's own training branch and autograd (build container only):

    python tests/golden/make_train_golden.py

The V-generalised reference is put in train() mode and `VANeRF.batch_render_pifu_nerf`
(src/model.py:1103-1422) runs its training paths: random patch (:1172-1189), stratified jitter (:1226-1230), view dropout
(:804-810), density noise (:1155-1156), random importance samples (:1439-1442).  torch / numpy global generators are seeded, so
`vanerf_b200.train.TrainRandom(seed, np_seed)` reproduces the same draws.  Loss = L1(tex_fg, 0.5) + 10 L1(tex_fg_fine, 0.5),
`loss.backward()` through the reference.  Stored: patch pixels, depths, colours, loss, and the gradient of every render-path
parameter (full tensor when small, norm + sum otherwise) and of the feature maps (norm + sum)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

H, W, V, PATCH, S_C, S_F = 512, 334, 3, 6, 16, 16
SEED, NP_SEED = 1234, 77


def patch_mask():
    m = np.zeros((1, H, W), bool)
    m[0, 230:290, 140:200] = True          # around the hands in the target view
    return m


def main():
    import torch
    from oracle import ref_import
    from vanerf_b200 import synthetic, weights
    ns = ref_import.load(patched=True)
    M = ns.model
    inp = synthetic.to_torch(synthetic.make_scene(H, W, V))
    sd = weights.init_state_dict(H, W, mode="stress")
    net = ref_import.build_net(ns, H, W, weights.to_torch(sd))
    net.train()
    net.train_out_h = net.train_out_w = PATCH
    for p in net.parameters():
        p.requires_grad_(True)
    fg = [t.clone().requires_grad_(True) for t in inp["feat_geo"]]
    ft = inp["feat_tex"].clone().requires_grad_(True)
    zs = []
    orr = M.VANeRF.rgba2out

    def r(self_, rgba, z, sdf):
        zs.append(z.detach().clone())
        return orr(self_, rgba, z, sdf)
    M.VANeRF.rgba2out = staticmethod(r)
    torch.manual_seed(SEED)
    np.random.seed(NP_SEED)
    out = M.VANeRF.batch_render_pifu_nerf(
        net, inp['img'], inp['cam_in'], inp['hand_type'], inp['targets'], V, inp['cam_tar'], 1, torch.zeros(1, 2), None, fg, ft, None,
        dict(inp['sp_data']), inp['objcenter'], fine=True, uniform=False, sample_per_ray_c=S_C, sample_per_ray_f=S_F, rand_noise_std=0.01,
        src_foreground_mask=inp['src_foreground_mask'], bounds=inp['bounds'], msk=torch.from_numpy(patch_mask()))
    loss = (out["tex_fg"] - 0.5).abs().mean() + 10.0 * (out["tex_fg_fine"] - 0.5).abs().mean()
    loss.backward()
    f = lambda t: t.detach().cpu().numpy()
    g = dict(H=H, W=W, V=V, patch=PATCH, S_c=S_C, S_f=S_F, seed=SEED, np_seed=NP_SEED, loss=f(loss),
             z=f(zs[0])[0], z_fine=f(zs[1])[0], tex_fg=f(out["tex_fg"])[0].reshape(3, -1).T, tex_fg_fine=f(out["tex_fg_fine"])[0].reshape(3, -1).T,
             alpha_fine=f(out["alpha_fine"])[0].reshape(-1))
    names = []
    for k, p in net.named_parameters():
        if k.startswith(("geo_encoder", "tex_encoder", "vgg_loss", "sp_encoder")) or p.grad is None:
            continue
        names.append(k)
        gr = f(p.grad).astype(np.float64)
        g["gn:" + k] = np.array([np.sqrt((gr ** 2).sum()), gr.sum()])
        if gr.size <= 20000:
            g["g:" + k] = gr.astype(np.float32)
    g["param_names"] = np.array(names)
    for nm, t in (("feat_geo0", fg[0]), ("feat_geo1", fg[1]), ("feat_tex", ft)):
        gr = f(t.grad).astype(np.float64)
        g["gn:" + nm] = np.array([np.sqrt((gr ** 2).sum()), gr.sum()])
    np.savez_compressed(os.path.join(HERE, "train_v3.npz"), **g)
    print("loss", float(loss), "params with grad", len(names), "z", g["z"].shape, "z_fine", g["z_fine"].shape)
    print({k: g["gn:" + k][0] for k in names[:8]})


if __name__ == "__main__":
    main()
