"""CPU: the CLIP oracle's preprocessing against the golden written by oracle/pin_clip_against_hf.py (the full 24-layer
forward is exercised by the pinning script itself: ~25 s of CPU, too slow for this suite) and the weight recipe."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_preprocess_matches_golden_and_state_has_the_reference_layout():
    from o2345 import synthetic as S
    from oracle import clip_oracle as CO
    from oracle.pin_clip_against_hf import clip_input
    gold = np.load(os.path.join(ROOT, "tests", "golden", "clip_mini.npz"))
    x224 = CO.preprocess(torch.from_numpy(clip_input()))
    assert float((x224[:, :, ::8, ::8] - torch.from_numpy(gold["x224"])).abs().max()) < 1e-6
    assert gold["embed"].shape == (1, 768)
    import torch as T
    from o2345.clip_image import FrozenCLIPImageEmbedder
    with T.device("meta"):
        keys = FrozenCLIPImageEmbedder().state_dict()
    assert "model.visual.transformer.resblocks.23.attn.in_proj_weight" in keys and "model.visual.proj" in keys
    assert sum(v.numel() for v in keys.values()) == 303_966_208          # OpenAI CLIP ViT-L/14 vision tower
