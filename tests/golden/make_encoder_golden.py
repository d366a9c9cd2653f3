"""Generates tests/golden/encoders.npz from the REFERENCE's own CNN encoders (build container only):

    python tests/golden/make_encoder_golden.py

`HGFilterV2` and `ResBlkEncoder` are imported from /root/reference/src/utils.py through oracle/ref_import.py, constructed with the
`configs/vanerf.json` arguments, loaded with name-keyed synthetic weights (`vanerf_b200.encoders.seeded_state_dict`, applied to the
reference modules' own state_dict keys), and run (CPU, fp32) on a seeded 2 x 3 x 128 x 128 image through the arithmetic of
`VANeRF.attach_geo_feat` / `attach_tex_feat` (src/model.py:711-738: average pooling `ds` times, 2 x - 1).  Stored: the input, the
reference's state_dict keys and shapes (28.3 M parameters: the test regenerates the weights from the names) and the three output maps."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

GEO_ARGS = {"n_stack": 1, "n_downsample": 4, "out_ch": 64, "hd": False}                                        # configs/vanerf.json:56-61
TEX_ARGS = {"ngf": 64, "n_downsample": 3, "n_blocks": 4, "n_upsample": 2, "out_ch": 8, "norm": "instance"}      # :92-99
SEED, V, H, W = 1234, 2, 128, 128


def main():
    import torch
    import torch.nn.functional as thf
    from oracle import ref_import
    ns = ref_import.load(patched=True)
    from vanerf_b200.encoders import seeded_state_dict
    geo = ns.utils.HGFilterV2(**GEO_ARGS).eval()
    tex = ns.utils.ResBlkEncoder(**TEX_ARGS).eval()
    geo.load_state_dict(seeded_state_dict(geo, SEED), strict=True)
    tex.load_state_dict(seeded_state_dict(tex, SEED), strict=True)
    g = torch.Generator().manual_seed(SEED + 1)
    im = torch.rand(V, 3, H, W, generator=g)
    with torch.no_grad():
        ds = thf.avg_pool2d(im, 2, stride=2)
        g0, g1 = geo(2.0 * ds - 1.0)
        t = tex(2.0 * ds - 1.0)
    out = {"im": im.numpy(), "geo0": g0.numpy(), "geo1": g1.numpy(), "tex": t.numpy()}
    sd = {**{"geo_encoder." + k: v for k, v in geo.state_dict().items()}, **{"tex_encoder." + k: v for k, v in tex.state_dict().items()}}
    out["keys"] = np.array(sorted(sd.keys()))
    out["shapes"] = np.array([str(tuple(sd[k].shape)) for k in sorted(sd.keys())])
    np.savez_compressed(os.path.join(HERE, "encoders.npz"), **out)
    print({k: getattr(v, "shape", v) for k, v in out.items()}, sum(v.numel() for v in sd.values()), "parameters")


if __name__ == "__main__":
    main()
