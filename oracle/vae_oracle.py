"""CPU oracle for the VAE (rows A6/A7): fp32 torch restatement of reference ldm/modules/diffusionmodules/model.py
:82-141 (ResnetBlock), :178-202 (AttnBlock), :434-459 (Encoder.forward), :535-568 (Decoder.forward) and
ldm/models/autoencoder.py:324-333, driven by a state dict with the reference's keys.  TEST INFRASTRUCTURE."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gn(x, sd, p):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _conv(x, sd, p, stride=1, pad=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride, pad)


def resnet_block(x, sd, p):
    h = _conv(F.silu(_gn(x, sd, p + ".norm1")), sd, p + ".conv1")
    h = _conv(F.silu(_gn(h, sd, p + ".norm2")), sd, p + ".conv2")
    if p + ".nin_shortcut.weight" in sd:
        x = _conv(x, sd, p + ".nin_shortcut", pad=0)
    return x + h


def attn_block(x, sd, p):
    B, C, H, W = x.shape
    h = _gn(x, sd, p + ".norm")
    q, k, v = (_conv(h, sd, f"{p}.{n}", pad=0).reshape(B, C, H * W) for n in "qkv")
    w = torch.bmm(q.permute(0, 2, 1), k) * int(C) ** -0.5
    w = F.softmax(w, dim=2)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(B, C, H, W)
    return x + _conv(h, sd, p + ".proj_out", pad=0)


def decode(sd, z, num_resolutions=4, num_res_blocks=2):
    h = _conv(_conv(z, sd, "post_quant_conv", pad=0), sd, "decoder.conv_in")
    h = resnet_block(h, sd, "decoder.mid.block_1")
    h = attn_block(h, sd, "decoder.mid.attn_1")
    h = resnet_block(h, sd, "decoder.mid.block_2")
    for lvl in reversed(range(num_resolutions)):
        for b in range(num_res_blocks + 1):
            h = resnet_block(h, sd, f"decoder.up.{lvl}.block.{b}")
        if lvl != 0:
            h = _conv(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd, f"decoder.up.{lvl}.upsample.conv")
    return _conv(F.silu(_gn(h, sd, "decoder.norm_out")), sd, "decoder.conv_out")


def encode_moments(sd, x, num_resolutions=4, num_res_blocks=2):
    h = _conv(x, sd, "encoder.conv_in")
    for lvl in range(num_resolutions):
        for b in range(num_res_blocks):
            h = resnet_block(h, sd, f"encoder.down.{lvl}.block.{b}")
        if lvl != num_resolutions - 1:
            h = _conv(F.pad(h, (0, 1, 0, 1)), sd, f"encoder.down.{lvl}.downsample.conv", stride=2, pad=0)
    h = resnet_block(h, sd, "encoder.mid.block_1")
    h = attn_block(h, sd, "encoder.mid.attn_1")
    h = resnet_block(h, sd, "encoder.mid.block_2")
    h = _conv(F.silu(_gn(h, sd, "encoder.norm_out")), sd, "encoder.conv_out")
    return _conv(h, sd, "quant_conv", pad=0)
