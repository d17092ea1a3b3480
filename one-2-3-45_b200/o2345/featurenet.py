"""FeatureNet + fused pyramid maps on the o2345 CUDA kernels.

Mirror of reference reconstruction/models/featurenet.py:12-91 (same constructor, same state-dict
keys, same return value of forward()) and of GenericTrainer.obtain_pyramid_feature_maps
(reference trainer_generic.py:1104-1125).  Every conv / InPlaceABN / bilinear up-sampling runs
through libo2345_sm100.so; InPlaceABN uses batch statistics because the reference never
switches these modules to eval mode.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._lib import inference_only
from . import ops


class ABNParams(nn.Module):
    """Parameter holder with InPlaceABN's state-dict keys (weight, bias, running_mean, running_var)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.eps, self.slope = 1e-5, 0.01


class ConvBnReLU(nn.Module):
    """conv (no bias) -> InPlaceABN(leaky_relu 0.01), reference featurenet.py:12-22."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1, norm_act=None):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = ABNParams(out_channels)
        self.stride, self.pad = stride, pad

    def run(self, x, out=None, layout="nchw", c0=0):
        """x NCHW fp32 -> normalised output written into `out` (allocated NCHW if None)."""
        raw, stats = ops.conv2d(ops.cf32(x), ops.cf32(self.conv.weight), None, self.stride, self.pad, True)
        if out is None:
            out = torch.empty_like(raw)
        ops.abn_apply(raw, stats, ops.cf32(self.bn.weight), ops.cf32(self.bn.bias), ops.view_of(out, layout, c0),
                      self.bn.eps, self.bn.slope)
        return out

    def forward(self, x):
        return self.run(x)


def _plain_conv(x, conv):
    out, _ = ops.conv2d(ops.cf32(x), ops.cf32(conv.weight), ops.cf32(conv.bias), 1, conv.padding[0], False)
    return out


class FeatureNet(nn.Module):
    """FPN that outputs 3 levels of features [feat2 (H/4), feat1 (H/2), feat0 (H)]."""

    def __init__(self, norm_act=None):
        super().__init__()
        self.conv0 = nn.Sequential(ConvBnReLU(3, 8, 3, 1, 1), ConvBnReLU(8, 8, 3, 1, 1))
        self.conv1 = nn.Sequential(ConvBnReLU(8, 16, 5, 2, 2), ConvBnReLU(16, 16, 3, 1, 1), ConvBnReLU(16, 16, 3, 1, 1))
        self.conv2 = nn.Sequential(ConvBnReLU(16, 32, 5, 2, 2), ConvBnReLU(32, 32, 3, 1, 1), ConvBnReLU(32, 32, 3, 1, 1))
        self.toplayer = nn.Conv2d(32, 32, 1)
        self.lat1 = nn.Conv2d(16, 32, 1)
        self.lat0 = nn.Conv2d(8, 32, 1)
        self.smooth1 = nn.Conv2d(32, 16, 3, padding=1)
        self.smooth0 = nn.Conv2d(32, 8, 3, padding=1)

    @inference_only
    def forward(self, x):
        x = ops.cf32(x)
        conv0 = self.conv0(x)
        conv1 = self.conv1(conv0)
        conv2 = self.conv2(conv1)
        feat2 = _plain_conv(conv2, self.toplayer)
        f1 = conv1.new_empty(conv1.shape[0], 32, conv1.shape[2], conv1.shape[3])
        ops.upsample_bilinear(feat2, 2, ops.view_of(f1, "nchw"), add=_plain_conv(conv1, self.lat1))
        f0 = conv0.new_empty(conv0.shape[0], 32, conv0.shape[2], conv0.shape[3])
        ops.upsample_bilinear(f1, 2, ops.view_of(f0, "nchw"), add=_plain_conv(conv0, self.lat0))
        feat1 = _plain_conv(f1, self.smooth1)
        feat0 = _plain_conv(f0, self.smooth0)
        return [feat2, feat1, feat0]


@torch.no_grad()
def obtain_pyramid_feature_maps(extractor: FeatureNet, imgs):
    """[V,3,H,W] -> fused [V,56,H,W] = cat(up4(feat2), up2(feat1), feat0), written in place by the
    up-sampling kernels (no torch.cat pass)."""
    feat2, feat1, feat0 = extractor(imgs)
    V, _, H, W = feat0.shape
    fused = torch.empty(V, 56, H, W, dtype=torch.float32, device=feat0.device)
    ops.upsample_bilinear(feat2, 4, ops.view_of(fused, "nchw", 0))
    ops.upsample_bilinear(feat1, 2, ops.view_of(fused, "nchw", 32))
    ops.upsample_bilinear(feat0, 1, ops.view_of(fused, "nchw", 48))
    return fused


@torch.no_grad()
def source_maps_channel_last(feature_maps, color_maps):
    """[V,56,H,W] + [V,3,H,W] -> channel-last [V,H,W,60] = (rgb, features, 0) for the blend kernel."""
    V, _, H, W = feature_maps.shape
    maps = torch.zeros(V, H, W, 60, dtype=torch.float32, device=feature_maps.device)
    ops.upsample_bilinear(ops.cf32(color_maps), 1, ops.view_of(maps, "nhwc", 0))
    ops.upsample_bilinear(ops.cf32(feature_maps), 1, ops.view_of(maps, "nhwc", 3))
    return maps
