"""Source-view CNN encoders: the step in front of the render path.

    HGFilterV2      src/utils.py:456-554   geometry encoder: stacked hourglass, GroupNorm  -> [geo0 (64, H/8, W/8), geo1 (8, H/2, W/2)]
    ConvBlock       src/utils.py:556-617   pre-activation residual block with a 1/2 + 1/4 + 1/4 channel split
    HourGlass       src/utils.py:400-443   recursive down / up branch, bicubic x2 up-sampling (align_corners)
    ResBlkEncoder   src/utils.py:353-398   texture encoder: 7x7 stem, strided convs, residual blocks, transposed convs, InstanceNorm
    attach_geo_feat / attach_tex_feat      src/model.py:711-738  (average-pool `ds` times, map [0,1] -> [-1,1], encode)

These are plain convolution stacks, i.e. library work (cuDNN through torch.nn), not part of the hand-written hot path; they are
here so that a user of the reference can go from images to a rendered view without leaving this package.  What is specific to this
implementation:

* parameter names are the reference's (`conv1.weight`, `m0.b2_3.bn1.weight`, `layers.13.layers.5.bias`, ...) so a reference
  checkpoint's `geo_encoder.*` / `tex_encoder.*` entries load with `strict=True`;
* the modules are built from shape tables instead of hand-unrolled constructors, run in `channels_last` (NHWC, the layout the
  gather kernels read and cuDNN's tensor-core kernels want) and optionally under bf16 autocast;
* `encode_geo` accepts any image size: the hourglass needs (H/8, W/8) divisible by 2^depth and the reference simply fails
  otherwise (e.g. for 512 x 334, its own data format); here the pooled image is replicate-padded on the right / bottom to the next
  multiple of 8 * 2^depth and the maps are cropped back to ceil(H/8) x ceil(W/8) and H/2 x W/2 (for sizes the reference supports
  nothing is padded and the results are the reference's).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def _gn(ch: int) -> nn.GroupNorm:
    return nn.GroupNorm(min(32, ch), ch)


class ConvBlock(nn.Module):
    """Pre-activation block: three 3x3 convolutions of out/2, out/4, out/4 channels, each fed by GroupNorm + ReLU of the previous
    result, concatenated and added to the (projected) input."""

    def __init__(self, in_planes: int, out_planes: int, norm: str = "group"):
        super().__init__()
        if norm != "group":
            raise NotImplementedError("only the GroupNorm flavour of the configs is built")
        widths = (in_planes, out_planes // 2, out_planes // 4, out_planes // 4)
        for i in range(3):
            setattr(self, f"bn{i + 1}", _gn(widths[i]))
            setattr(self, f"conv{i + 1}", nn.Conv2d(widths[i], widths[i + 1], 3, 1, 1, bias=False))
        self.bn4 = _gn(in_planes)            # the reference registers it whether or not the projection exists
        self.downsample = None
        if in_planes != out_planes:
            self.downsample = nn.Sequential(self.bn4, nn.ReLU(), nn.Conv2d(in_planes, out_planes, 1, bias=False))

    def forward(self, x):
        parts, y = [], x
        for i in (1, 2, 3):
            y = getattr(self, f"conv{i}")(F.relu(getattr(self, f"bn{i}")(y)))
            parts.append(y)
        return torch.cat(parts, 1) + (x if self.downsample is None else self.downsample(x))


class HourGlass(nn.Module):
    def __init__(self, depth: int, num_features: int, norm: str = "group"):
        super().__init__()
        self.depth = depth
        for level in range(depth, 0, -1):
            for tag in ("b1_", "b2_", "b3_"):
                self.add_module(f"{tag}{level}", ConvBlock(num_features, num_features, norm))
        self.add_module("b2_plus_1", ConvBlock(num_features, num_features, norm))

    def _level(self, level: int, x):
        skip = self._modules[f"b1_{level}"](x)
        low = self._modules[f"b2_{level}"](F.avg_pool2d(x, 2, stride=2))
        low = self._level(level - 1, low) if level > 1 else self._modules["b2_plus_1"](low)
        low = self._modules[f"b3_{level}"](low)
        return skip + F.interpolate(low, scale_factor=2, mode="bicubic", align_corners=True)

    def forward(self, x):
        return self._level(self.depth, x)


class DeconvReLUGroup(nn.Module):
    def __init__(self, in_ch: int, out_ch: int, bias: bool = False):
        super().__init__()
        self.conv = nn.ConvTranspose2d(in_ch, out_ch, 3, stride=2, padding=1, output_padding=1, bias=bias)
        self.norm = _gn(out_ch)

    def forward(self, x):
        return F.relu(self.norm(self.conv(x)))


class HGFilterV2(nn.Module):
    def __init__(self, in_ch: int = 3, out_ch: int = 128, n_stack: int = 2, n_downsample: int = 4, norm: str = "group", hd: bool = False,
                 **kwargs):
        super().__init__()
        self.n_stack, self.hd, self.depth = n_stack, hd, n_downsample
        self.unpack1 = DeconvReLUGroup(128, 32)
        self.conv_out = nn.Conv2d(32, kwargs.get("out_ch_hd", 8), 5, padding=2)
        self.conv1 = nn.Conv2d(in_ch, 64, 7, stride=2, padding=3)
        self.bn1 = nn.GroupNorm(32, 64)
        for name, (ci, co) in (("conv2", (64, 128)), ("conv3", (128, 128)), ("conv4", (128, 256))):
            self.add_module(name, ConvBlock(ci, co, norm))
        for i in range(n_stack):
            self.add_module(f"m{i}", HourGlass(n_downsample, 256, norm))
            self.add_module(f"top_m_{i}", ConvBlock(256, 256, norm))
            self.add_module(f"conv_last{i}", nn.Conv2d(256, 256, 1))
            self.add_module(f"bn_end{i}", nn.GroupNorm(32, 256))
            self.add_module(f"l{i}", nn.Conv2d(256, out_ch, 1))
            if i + 1 < n_stack:
                self.add_module(f"bl{i}", nn.Conv2d(256, 256, 1))
                self.add_module(f"al{i}", nn.Conv2d(out_ch, 256, 1))

    @property
    def size_multiple(self) -> int:
        """Input sizes must be multiples of this for the hourglass skips to line up (stem /2, pool /2 unless hd, 2^depth)."""
        return (2 if self.hd else 4) * (1 << self.depth)

    def forward(self, x) -> List[torch.Tensor]:
        m = self._modules
        x = self.conv2(F.relu(self.bn1(self.conv1(x))))
        x_hd = self.conv_out(self.unpack1(x))
        if not self.hd:
            x = F.avg_pool2d(x, 2, stride=2)
        prev = self.conv4(self.conv3(x))
        out = None
        for i in range(self.n_stack):
            y = m[f"top_m_{i}"](m[f"m{i}"](prev))
            y = F.relu(m[f"bn_end{i}"](m[f"conv_last{i}"](y)))
            out = m[f"l{i}"](y)
            if i + 1 < self.n_stack:
                prev = prev + m[f"bl{i}"](y) + m[f"al{i}"](out)
        return [out, x_hd]


class _ResBlk(nn.Module):
    def __init__(self, ch: int, norm_layer):
        super().__init__()
        self.layers = nn.Sequential(nn.ReplicationPad2d(1), nn.Conv2d(ch, ch, 3), norm_layer(ch), nn.ReLU(),
                                    nn.ReplicationPad2d(1), nn.Conv2d(ch, ch, 3), norm_layer(ch))

    def forward(self, x):
        return x + self.layers(x)


class ResBlkEncoder(nn.Module):
    def __init__(self, in_ch: int = 3, out_ch: int = 8, ngf: int = 16, n_downsample: int = 3, n_blocks: int = 4, n_upsample: int = 3,
                 norm: str = "instance"):
        super().__init__()
        if norm != "instance":
            raise NotImplementedError("only the InstanceNorm flavour of the configs is built")
        norm_layer = lambda ch: nn.InstanceNorm2d(ch, affine=False, track_running_stats=False)
        # the Sequential's indices are the parameter names of a reference checkpoint: keep one entry per reference layer
        seq: list = [nn.ReplicationPad2d(3), nn.Conv2d(in_ch, ngf, 7), norm_layer(ngf), nn.ReLU()]
        ch = ngf
        for _ in range(n_downsample):
            seq += [nn.Conv2d(ch, 2 * ch, 3, stride=2, padding=1), norm_layer(2 * ch), nn.ReLU()]
            ch *= 2
        seq += [_ResBlk(ch, norm_layer) for _ in range(n_blocks)]
        for _ in range(n_upsample):
            seq += [nn.ConvTranspose2d(ch, ch // 2, 3, stride=2, padding=1, output_padding=1), norm_layer(ch // 2), nn.ReLU()]
            ch //= 2
        if n_upsample > 0:
            seq += [nn.ReplicationPad2d(3), nn.Conv2d(ch, out_ch, 7)]
        self.layers = nn.Sequential(*seq)

    def forward(self, x):
        return self.layers(x)


# ---------------------------------------------------------------------------------------------------- attach_*_feat
def _pool_and_center(im: torch.Tensor, ds: int) -> torch.Tensor:
    if im.dim() == 5:
        im = im.reshape(-1, *im.shape[2:])
    for _ in range(ds):
        im = F.avg_pool2d(im, 2, stride=2)
    return 2.0 * im - 1.0


class Encoders(nn.Module):
    """Both encoders with the reference's attribute names (`geo_encoder`, `tex_encoder`), so that `load_state_dict` of a reference
    checkpoint (strict=False: the render-path keys belong to the fused kernels) fills them."""

    def __init__(self, geo_args: Optional[dict] = None, tex_args: Optional[dict] = None, ds_geo: int = 1, ds_tex: int = 1):
        super().__init__()
        # configs/vanerf.json:40-41,56-61,92-99
        self.geo_encoder = HGFilterV2(**(geo_args or {"n_stack": 1, "n_downsample": 4, "out_ch": 64, "hd": False}))
        self.tex_encoder = ResBlkEncoder(**(tex_args or {"ngf": 64, "n_downsample": 3, "n_blocks": 4, "n_upsample": 2, "out_ch": 8,
                                                         "norm": "instance"}))
        self.ds_geo, self.ds_tex = ds_geo, ds_tex
        self.autocast_dtype: Optional[torch.dtype] = None          # torch.bfloat16: tensor-core convolutions (maps stay fp32)

    def _run(self, net, x):
        x = x.contiguous(memory_format=torch.channels_last)
        if self.autocast_dtype is not None and x.is_cuda:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                y = net(x)
        else:
            y = net(x)
        return y

    @torch.no_grad()
    def encode_geo(self, im: torch.Tensor) -> List[torch.Tensor]:
        """im (V,3,H,W) or (B,V,3,H,W) in [0,1] -> [geo0 (V,C,ceil(H'/4),ceil(W'/4)), geo1 (V,8,H',W')], H' = H / 2^ds_geo."""
        x = _pool_and_center(im, self.ds_geo)
        h, w = x.shape[-2:]
        m = self.geo_encoder.size_multiple
        ph, pw = (-h) % m, (-w) % m
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph), mode="replicate")
        g0, g1 = self._run(self.geo_encoder, x)
        g0 = g0[..., : -(-h // 4), : -(-w // 4)] if not self.geo_encoder.hd else g0[..., : -(-h // 2), : -(-w // 2)]
        return [g0.float().contiguous(), g1[..., :h, :w].float().contiguous()]

    @torch.no_grad()
    def encode_tex(self, im: torch.Tensor) -> torch.Tensor:
        return self._run(self.tex_encoder, _pool_and_center(im, self.ds_tex)).float().contiguous()


def seeded_state_dict(module: nn.Module, seed: int = 0) -> dict:
    """Deterministic synthetic weights keyed by PARAMETER NAME (not by construction order), so that any implementation with the
    reference's state_dict keys - this module or the reference's own classes - gets bit-identical tensors: weights ~ N(0, 1/fan_in),
    norm scales 1 + 0.1 N(0,1), biases 0.1 N(0,1).  Aliased entries (`bn4` is also `downsample.0`) get the same values."""
    import zlib
    out = {}
    for key, ref in module.state_dict().items():
        canon = key.replace(".downsample.0.", ".bn4.")
        g = torch.Generator().manual_seed((zlib.crc32(canon.encode()) + 7919 * seed) & 0x7FFFFFFF)
        x = torch.randn(ref.shape, generator=g, dtype=torch.float32)
        if ref.dim() >= 2:
            x = x / float(ref[0].numel()) ** 0.5
        elif key.endswith("weight"):
            x = 1.0 + 0.1 * x
        else:
            x = 0.1 * x
        out[key] = x
    return out
