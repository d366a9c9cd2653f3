"""Seeded inputs of the stage-level parity checks (tests/test_stages.py, tests/golden/make_stage_golden.py): reference
layouts, B = 1, V = 3 source views, N = 96 query samples, 64 x 64 source images (the LayerNorm shapes of TexVisFusion follow
the map sizes)."""
import numpy as np

V, N, NV, H, W = 3, 96, 1558, 64, 64


def make(seed=7):
    r = np.random.RandomState(seed)
    f = lambda *s: r.standard_normal(s).astype(np.float32)
    u = lambda lo, hi, *s: r.uniform(lo, hi, s).astype(np.float32)
    d = dict(
        vert_xy=u(-1.1, 1.1, V, NV, 2), g0=f(V, 64, H // 8, W // 8), g1=f(V, 8, H // 2, W // 2), tex=f(V, 8, H // 4, W // 4), img=u(0, 1, V, 3, H, W),
        px64=f(V, N, 64), px8=f(V, N, 8), vert=np.repeat(u(-0.1, 0.1, 1, NV, 3), V, 0), v=np.repeat(u(-0.12, 0.12, 1, N, 3), V, 0),
        vert_vis=(r.rand(V, NV, 1) > 0.4).astype(np.float32), query_vis=(r.rand(V, N, 1) > 0.4).astype(np.float32),
        query_sdf=np.repeat(u(-0.02, 0.05, 1, N, 1), V, 0), ft_xy=f(V, N, 8), img_xy=u(0, 1, V, N, 3), latent24=f(V, N, 24),
        uv=u(-1.2, 1.2, V, N, 2),
        pe=0.3 * f(1, V, N, 294), f64=f(1, V, N, 64), f8=f(1, V, N, 8), a=(r.rand(1, 1, N, 1) > 0.2).astype(np.float32).repeat(V, 1),
        rgb_feats=f(8, 12, V, 40), ray_diffs=np.concatenate([f(8, 12, V, 3), u(0.5, 1.0, 8, 12, V, 1)], -1),
        proj_mask=(r.rand(8, 12, 1, 1) > 0.2).astype(np.float32).repeat(V, 2),
        extrin=np.tile(np.eye(4, dtype=np.float32), (V, 1, 1)), kpt3d=u(-0.1, 0.1, 1, 42, 3),
    )
    for i in range(V):                                   # distinct source cameras: small rotations about y + translation
        c, s = np.cos(0.3 * i), np.sin(0.3 * i)
        d["extrin"][i, :3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32)
        d["extrin"][i, :3, 3] = np.array([0.01 * i, -0.02, 1.0], np.float32)
    w = r.rand(1, V, N, 1).astype(np.float32) * d["a"]
    d["w"] = (w / (w.sum(1, keepdims=True) + 1e-6)).astype(np.float32)
    return d
