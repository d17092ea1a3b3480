"""CPU oracle for path A: Zero123 UNet forward, DDIM schedule and the CFG + DDIM update.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE (same rules as recon_oracle.py).  fp32 torch-CPU restatement of
  ldm/modules/diffusionmodules/openaimodel.py:63-276,414-777, ldm/modules/attention.py:37-64,152-266,
  ldm/modules/diffusionmodules/util.py:21-74,151-171, ldm/models/diffusion/ddim.py:37-243,
driven by a plain state dict with the reference's keys.  Pinned against the reference's own UNetModel and
DDIMSampler by oracle/pin_ldm_against_reference.py (golden vectors in tests/golden/ldm_mini.npz).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

CHANNEL_MULT = (1, 2, 4, 4)
ATTN_DS = (1, 2, 4)


def timestep_embedding(t, dim):
    """reference util.py:151-171"""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], -1)


def _gn(x, sd, p, eps=1e-5):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _conv(x, sd, p, stride=1, pad=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride, pad)


def _lin(x, sd, p):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def res_block(x, emb, sd, p):
    """reference openaimodel.py:256-276 (no up/down, no scale-shift norm)"""
    h = _conv(F.silu(_gn(x, sd, p + ".in_layers.0")), sd, p + ".in_layers.2")
    h = h + _lin(F.silu(emb), sd, p + ".emb_layers.1")[:, :, None, None]
    h = _conv(F.silu(_gn(h, sd, p + ".out_layers.0")), sd, p + ".out_layers.3")
    skip = _conv(x, sd, p + ".skip_connection", pad=0) if (p + ".skip_connection.weight") in sd else x
    return skip + h


def attention(x, ctx, sd, p, heads):
    """reference attention.py:170-193"""
    q, k, v = _lin(x, sd, p + ".to_q"), _lin(ctx, sd, p + ".to_k"), _lin(ctx, sd, p + ".to_v")
    B, N, C = q.shape
    d = C // heads
    sp = lambda t: t.view(B, -1, heads, d).permute(0, 2, 1, 3)
    s = torch.einsum("bhid,bhjd->bhij", sp(q), sp(k)) * d ** -0.5
    o = torch.einsum("bhij,bhjd->bhid", s.softmax(-1), sp(v)).permute(0, 2, 1, 3).reshape(B, N, C)
    return _lin(o, sd, p + ".to_out.0")


def spatial_transformer(x, ctx, sd, p, heads):
    """reference attention.py:255-266, 214-218, 37-64"""
    B, C, H, W = x.shape
    h = _conv(_gn(x, sd, p + ".norm", 1e-6), sd, p + ".proj_in", pad=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = p + ".transformer_blocks.0"
    ln = lambda t, n: F.layer_norm(t, (C,), sd[f"{b}.{n}.weight"], sd[f"{b}.{n}.bias"])
    n1 = ln(h, "norm1")
    h = attention(n1, n1, sd, b + ".attn1", heads) + h
    h = attention(ln(h, "norm2"), ctx, sd, b + ".attn2", heads) + h
    g = _lin(ln(h, "norm3"), sd, b + ".ff.net.0.proj")
    a, gate = g.chunk(2, -1)
    h = _lin(a * F.gelu(gate), sd, b + ".ff.net.2") + h
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _conv(h, sd, p + ".proj_out", pad=0) + x


def unet_plan(num_res_blocks=2):
    """Layer kinds per block, in the order the reference builds them (openaimodel.py:535-712)."""
    inp, out, ds = [["conv"]], [], 1
    for level in range(4):
        for _ in range(num_res_blocks):
            inp.append(["res", "attn"] if ds in ATTN_DS else ["res"])
        if level != 3:
            inp.append(["down"])
            ds *= 2
    for level in range(3, -1, -1):
        for i in range(num_res_blocks + 1):
            l = ["res", "attn"] if ds in ATTN_DS else ["res"]
            if level and i == num_res_blocks:
                l.append("up")
                ds //= 2
            out.append(l)
    return inp, out


def unet_forward(sd, x, t, context, heads=8, model_channels=320):
    """UNetModel.forward (reference openaimodel.py:745-777) in fp32."""
    emb = _lin(F.silu(_lin(timestep_embedding(t, model_channels), sd, "time_embed.0")), sd, "time_embed.2")
    inp, out = unet_plan()

    def run(h, prefix, kinds):
        for j, kind in enumerate(kinds):
            p = f"{prefix}.{j}"
            if kind == "conv":
                h = _conv(h, sd, p)
            elif kind == "res":
                h = res_block(h, emb, sd, p)
            elif kind == "attn":
                h = spatial_transformer(h, context, sd, p, heads)
            elif kind == "down":
                h = _conv(h, sd, p + ".op", stride=2)
            elif kind == "up":
                h = _conv(F.interpolate(h, scale_factor=2, mode="nearest"), sd, p + ".conv")
        return h

    hs, h = [], x
    for i, kinds in enumerate(inp):
        h = run(h, f"input_blocks.{i}", kinds)
        hs.append(h)
    h = run(h, "middle_block", ["res", "attn", "res"])
    for i, kinds in enumerate(out):
        h = run(torch.cat([h, hs.pop()], 1), f"output_blocks.{i}", kinds)
    return _conv(F.silu(_gn(h, sd, "out.0")), sd, "out.2")


# ----------------------------------------------------------------------------- DDIM (rows A1, A9)
def linear_beta_alphas_cumprod(n=1000, start=0.00085, end=0.0120):
    """reference util.py:21-25 + ddpm.py:126-178: fp64 schedule, stored as fp32."""
    betas = np.linspace(start ** 0.5, end ** 0.5, n, dtype=np.float64) ** 2
    return np.cumprod(1.0 - betas, axis=0).astype(np.float32)


def ddim_schedule(alphas_cumprod, S, eta, n_train=1000):
    """make_ddim_timesteps('uniform') + make_ddim_sampling_parameters (reference util.py:46-74)."""
    c = n_train // S
    ts = np.asarray(list(range(0, n_train, c))) + 1
    ac = np.asarray(alphas_cumprod, np.float64) if not torch.is_tensor(alphas_cumprod) else alphas_cumprod.double().numpy()
    a = ac[ts]
    a_prev = np.asarray([ac[0]] + ac[ts[:-1]].tolist())
    sig = eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
    return ts, a, a_prev, sig


def ddim_sample(apply_model, x_T, cond, uncond, scale, alphas_cumprod, S, eta, noises):
    """DDIMSampler.ddim_sampling with classifier-free guidance (reference ddim.py:130-243).  `noises[i]` is the
    randn drawn at iteration i.  Quirk kept: timesteps[:-1] are used (t_start = -1 drops the last entry)."""
    ts, a, a_prev, sig = ddim_schedule(alphas_cumprod, S, eta)
    ts = ts[:-1]
    x = x_T
    for i, step in enumerate(np.flip(ts)):
        index = len(ts) - i - 1
        B = x.shape[0]
        t = torch.full((2 * B,), int(step), dtype=torch.long)
        c_in = {k: [torch.cat([uncond[k][0], cond[k][0]])] for k in cond}
        e = apply_model(torch.cat([x] * 2), t, c_in)
        eu, ec = e.chunk(2)
        e_t = eu + scale * (ec - eu)
        at, ap, st = float(a[index]), float(a_prev[index]), float(sig[index])
        p0 = (x - math.sqrt(1 - at) * e_t) / math.sqrt(at)
        x = math.sqrt(ap) * p0 + math.sqrt(1.0 - ap - st ** 2) * e_t + st * noises[i]
    return x


class ToyModel:
    """The attributes DDIMSampler reads from LatentDiffusion (ddim.py:17-19,40-46,194) around a cheap epsilon model."""

    def __init__(self):
        ac = torch.from_numpy(linear_beta_alphas_cumprod())
        self.num_timesteps, self.device = 1000, torch.device("cpu")
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = torch.cat([torch.ones(1), ac[:-1]])
        self.betas = 1 - ac / self.alphas_cumprod_prev
        g = torch.Generator().manual_seed(11)
        self.w = torch.randn(4, 8, 3, 3, generator=g) * 0.2

    def apply_model(self, x, t, c):
        xc = torch.cat([x, c["c_concat"][0]], 1)
        return torch.nn.functional.conv2d(xc, self.w, padding=1) * (1.0 + c["c_crossattn"][0].mean(dim=(1, 2))[:, None, None, None]) \
            + 0.001 * t.float()[:, None, None, None]

    def to(self, device):
        self.device = torch.device(device)
        for k in ("alphas_cumprod", "alphas_cumprod_prev", "betas", "w"):
            setattr(self, k, getattr(self, k).to(device))
        return self
