"""Pin oracle/ldm_oracle.py against the REAL reference (UNetModel, DDIMSampler) and freeze golden vectors.

Build container only:   python -m oracle.pin_ldm_against_reference
Writes tests/golden/ldm_mini.npz: the reference UNet's epsilon prediction for a seeded batch of 2 and a short
DDIM trajectory (S=5, eta=1, CFG scale 3) driven by the reference sampler with injected noise.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "one-2-3-45_b200"))
sys.path.insert(0, ROOT)

from o2345 import synthetic as S  # noqa: E402
from oracle import ldm_oracle as LO  # noqa: E402

REF = "/root/reference"


def unet_inputs(seed=3, batch=2):
    g = np.random.default_rng(seed)
    x = g.standard_normal((batch, 8, 32, 32), dtype=np.float32)
    t = np.array([981, 21][:batch], np.int64)
    ctx = g.standard_normal((batch, 1, 768), dtype=np.float32)
    return x, t, ctx


def import_reference_ldm():
    for name in ("matplotlib", "matplotlib.pyplot", "omegaconf", "omegaconf.listconfig", "taming", "kornia", "clip"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["omegaconf.listconfig"].ListConfig = list
    sys.modules["omegaconf"].ListConfig = list
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from ldm.modules.diffusionmodules.openaimodel import UNetModel
    from ldm.models.diffusion.ddim import DDIMSampler
    return UNetModel, DDIMSampler


ToyModel = LO.ToyModel


def main():
    UNetModel, DDIMSampler = import_reference_ldm()
    ok = True
    gold = {}
    sd = {k: torch.from_numpy(v) for k, v in S.unet_state(0).items()}
    ref = UNetModel(image_size=32, in_channels=8, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
                    num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
                    transformer_depth=1, context_dim=768, use_checkpoint=False, legacy=False)
    res = ref.load_state_dict(sd, strict=True)
    x, t, ctx = unet_inputs()
    with torch.no_grad():
        e_ref = ref(torch.from_numpy(x), torch.from_numpy(t), torch.from_numpy(ctx))
        e_o = LO.unet_forward(sd, torch.from_numpy(x), torch.from_numpy(t), torch.from_numpy(ctx))
    err = float((e_ref - e_o).abs().max())
    print(f"[{'ok ' if err < 2e-4 else 'BAD'}] A2-A4 UNet forward (fp32 CPU) max|ref-oracle| = {err:.3e} (scale {float(e_ref.abs().max()):.3g}, std {float(e_ref.std()):.3g})")
    ok &= err < 2e-4
    gold["unet_eps"] = e_ref.numpy()
    del ref

    # DDIM trajectory with the reference sampler
    toy = ToyModel()
    B, Sst, eta, scale = 2, 5, 1.0, 3.0
    g = np.random.default_rng(5)
    cond = {"c_crossattn": [torch.from_numpy(g.standard_normal((B, 1, 768), dtype=np.float32))],
            "c_concat": [torch.from_numpy(g.standard_normal((B, 4, 32, 32), dtype=np.float32))]}
    uc = {"c_crossattn": [torch.zeros(B, 1, 768)], "c_concat": [torch.zeros(B, 4, 32, 32)]}
    torch.manual_seed(123)
    x_T = torch.randn(B, 4, 32, 32)
    ts, a, a_prev, sig = LO.ddim_schedule(toy.alphas_cumprod, Sst, eta)
    noises = [torch.randn(B, 4, 32, 32) for _ in range(len(ts) - 1)]
    torch.manual_seed(123)
    _ = torch.randn(B, 4, 32, 32)          # the reference draws x_T itself when x_T is None; we pass it, so skip one draw
    sampler = DDIMSampler(toy)
    out_ref, _ = sampler.sample(S=Sst, batch_size=B, shape=[4, 32, 32], conditioning=cond, verbose=False, eta=eta, x_T=x_T,
                                unconditional_guidance_scale=scale, unconditional_conditioning=uc)
    out_o = LO.ddim_sample(toy.apply_model, x_T, cond, uc, scale, toy.alphas_cumprod, Sst, eta, noises)
    err = float((out_ref - out_o).abs().max())
    print(f"[{'ok ' if err < 1e-4 else 'BAD'}] A1/A9 DDIM trajectory ({len(ts) - 1} iterations for S={Sst}) max|ref-oracle| = {err:.3e}")
    ok &= err < 1e-4
    for S_ in (75, 50):
        n = len(LO.ddim_schedule(toy.alphas_cumprod, S_, 1.0)[0]) - 1
        print(f"      S={S_}: {n} iterations (reference notebook: {76 if S_ == 75 else 49})")
        ok &= n == (76 if S_ == 75 else 49)
    gold.update(ddim_out=out_ref.numpy(), ddim_alphas=np.asarray(sampler.ddim_alphas, np.float32),
                ddim_sigmas=np.asarray(sampler.ddim_sigmas, np.float32), ddim_timesteps=np.asarray(sampler.ddim_timesteps))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ldm_mini.npz"), **gold)
    print("ALL PINNED" if ok else "SOME CHECKS FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
