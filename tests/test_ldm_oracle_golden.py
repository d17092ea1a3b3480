"""CPU: the path-A oracle (UNet forward, DDIM schedule/update) reproduces the golden vectors frozen from the
reference's own UNetModel / DDIMSampler (tests/golden/ldm_mini.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import ldm_oracle as LO
from oracle.pin_ldm_against_reference import unet_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "ldm_mini.npz"))


def test_schedule_matches_reference(gold):
    toy = LO.ToyModel()
    ts, a, a_prev, sig = LO.ddim_schedule(toy.alphas_cumprod, 5, 1.0)
    assert np.array_equal(ts, gold["ddim_timesteps"])
    assert np.allclose(a, gold["ddim_alphas"], rtol=1e-6) and np.allclose(sig, gold["ddim_sigmas"], rtol=1e-5)
    assert len(LO.ddim_schedule(toy.alphas_cumprod, 75, 1.0)[0]) - 1 == 76      # 76 / 49 iterations
    assert len(LO.ddim_schedule(toy.alphas_cumprod, 50, 1.0)[0]) - 1 == 49


def test_ddim_trajectory(gold):
    toy = LO.ToyModel()
    B = 2
    g = np.random.default_rng(5)
    cond = {"c_crossattn": [torch.from_numpy(g.standard_normal((B, 1, 768), dtype=np.float32))],
            "c_concat": [torch.from_numpy(g.standard_normal((B, 4, 32, 32), dtype=np.float32))]}
    uc = {"c_crossattn": [torch.zeros(B, 1, 768)], "c_concat": [torch.zeros(B, 4, 32, 32)]}
    torch.manual_seed(123)
    x_T = torch.randn(B, 4, 32, 32)
    noises = [torch.randn(B, 4, 32, 32) for _ in range(4)]
    out = LO.ddim_sample(toy.apply_model, x_T, cond, uc, 3.0, toy.alphas_cumprod, 5, 1.0, noises)
    assert float((out - torch.from_numpy(gold["ddim_out"])).abs().max()) < 1e-4


def test_unet_forward(gold):
    from o2345 import synthetic as S
    sd = {k: torch.from_numpy(v) for k, v in S.unet_state(0).items()}
    x, t, ctx = unet_inputs()
    with torch.no_grad():
        e = LO.unet_forward(sd, torch.from_numpy(x), torch.from_numpy(t), torch.from_numpy(ctx))
    assert float((e - torch.from_numpy(gold["unet_eps"])).abs().max()) < 2e-4
