"""Per-frame TexVisFusion global vertex feature: the library's kernels (csrc/gfeat.cuh, vanerf_global_vertex_feature) against
the same stacks in torch fp32 (src/networks.py:246-279 semantics; cuDNN without TF32), on the bench-size and on odd-size maps."""
import numpy as np
import pytest
import torch

from vanerf_b200 import synthetic, weights
from vanerf_b200.renderer import Renderer

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,W,V,mode", [(512, 334, 3, "stress"), (256, 256, 1, "ref"), (70, 50, 2, "stress")])
def test_global_vertex_feature_kernels_match_torch(cuda_lib, H, W, V, mode):
    inp = synthetic.to_torch(synthetic.make_scene(H, W, V), "cuda:0")
    r = Renderer("cuda:0")
    r.load_state_dict(weights.init_state_dict(H, W, mode=mode))
    img, tex = inp["img"].float().contiguous(), inp["feat_tex"].float().contiguous()
    with torch.no_grad():
        got = r.global_vertex_feature(img, tex)
        with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=True, allow_tf32=False):
            ref = Renderer._global_vertex_feature(img, tex, r.sd)
    assert got.shape == ref.shape == (V, 1558, 18)
    err = float((got - ref).abs().max())
    scale = max(1.0, float(ref.abs().max()))
    print(f"gfeat {H}x{W} V={V} {mode}: max-abs err {err:.3e} (|ref| max {scale:.3f})")
    assert torch.isfinite(got).all() and err < 2e-4 * scale
