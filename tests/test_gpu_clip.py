"""GPU: the CLIP ViT-L/14 image tower (row A8) on the o2345 kernels against the golden embedding that
oracle/pin_clip_against_hf.py froze (Hugging Face transformers' implementation, fp32, same seeded weights).

Tolerance: 24 pre-LN transformer layers with fp16 activations (the reference runs the tower in fp16 under
--half_precision as well); the embedding has std 0.68 and |max| 2.5, the bar is 3x the measured deviation: max |err| < 1.2e-2, mean < 2.1e-3."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "clip_mini.npz"))


@pytest.fixture(scope="module")
def clip_net():
    from o2345 import synthetic as S
    from o2345.clip_image import FrozenCLIPImageEmbedder
    net = FrozenCLIPImageEmbedder()
    net.load_state_dict({k: torch.from_numpy(v) for k, v in S.clip_state().items()})
    return net.cuda().requires_grad_(False)


def test_preprocess_and_patch_gather(gold):
    """Bicubic (align_corners) 256 -> 224 resize + CLIP normalisation + 14 x 14 patch gather in one kernel."""
    from o2345 import ops_a
    from o2345.clip_image import CLIP_MEAN, CLIP_STD
    from oracle.pin_clip_against_hf import clip_input
    x = torch.from_numpy(clip_input()).cuda()
    p = ops_a.clip_patches(x, 224, 14, CLIP_MEAN, CLIP_STD, 592)
    assert p.shape == (256, 592) and float(p[:, 588:].abs().max()) == 0.0
    img = p[:, :588].float().view(16, 16, 3, 14, 14).permute(2, 0, 3, 1, 4).reshape(3, 224, 224)
    err = (img[:, ::8, ::8].cpu() - torch.from_numpy(gold["x224"][0])).abs().max()
    assert float(err) < 3e-3, float(err)                     # fp16 rounding of values up to 2.6


def test_embedding_matches_golden(clip_net, gold):
    from oracle.pin_clip_against_hf import clip_input
    x = torch.from_numpy(clip_input()).cuda()
    e = clip_net(x)
    torch.cuda.synchronize()
    assert e.shape == (1, 768) and e.dtype == torch.float32
    err = (e.cpu() - torch.from_numpy(gold["embed"])).abs()
    print("clip: max err", float(err.max()), "mean err", float(err.mean()))
    assert float(err.max()) < 1.2e-2 and float(err.mean()) < 2.1e-3     # 3x the measured 3.7e-3 / 7e-4
    c = clip_net.encode(torch.cat([x, -x]))                  # batch of two, encode() adds the token axis
    assert c.shape == (2, 1, 768) and float((c[0, 0].cpu() - torch.from_numpy(gold["embed"][0])).abs().max()) < 1.2e-2
    assert clip_net([""]).shape == (1, 768) and float(clip_net([""]).abs().max()) == 0.0
