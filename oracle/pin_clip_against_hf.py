"""Pin oracle/clip_oracle.py against Hugging Face transformers' CLIPVisionModelWithProjection (an independent
implementation of the OpenAI CLIP ViT-L/14 vision tower; the reference's own `clip` dependency is not available) with the
seeded synthetic weights, and freeze tests/golden/clip_mini.npz.   python -m oracle.pin_clip_against_hf"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "one-2-3-45_b200"))
sys.path.insert(0, ROOT)
from o2345 import synthetic as S  # noqa: E402
from oracle import clip_oracle as CO  # noqa: E402


def clip_input():
    """One 256 x 256 image in [-1, 1] (smooth + noise, so that the bicubic resize matters)."""
    g = np.random.default_rng(6)
    yy, xx = np.mgrid[0:256, 0:256].astype(np.float32) / 255.0
    base = np.stack([np.sin(6 * xx + 2 * yy), np.cos(5 * yy - xx), np.sin(9 * xx * yy)], 0)
    return np.clip(0.7 * base + 0.3 * g.standard_normal((3, 256, 256), dtype=np.float32), -1, 1)[None].astype(np.float32)


def to_hf(sd):
    """OpenAI CLIP `model.visual.*` names -> transformers CLIPVisionModelWithProjection names."""
    out, p = {}, "model.visual."
    out["vision_model.embeddings.class_embedding"] = sd[p + "class_embedding"]
    out["vision_model.embeddings.patch_embedding.weight"] = sd[p + "conv1.weight"]
    out["vision_model.embeddings.position_embedding.weight"] = sd[p + "positional_embedding"]
    out["vision_model.pre_layrnorm.weight"], out["vision_model.pre_layrnorm.bias"] = sd[p + "ln_pre.weight"], sd[p + "ln_pre.bias"]
    out["vision_model.post_layernorm.weight"], out["vision_model.post_layernorm.bias"] = sd[p + "ln_post.weight"], sd[p + "ln_post.bias"]
    out["visual_projection.weight"] = sd[p + "proj"].t()
    layers = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith(p + "transformer.resblocks."))
    for i in range(layers):
        a, b = f"{p}transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        w, bias = sd[a + "attn.in_proj_weight"].chunk(3, 0), sd[a + "attn.in_proj_bias"].chunk(3, 0)
        for n, ww, bb in zip("qkv", w, bias):
            out[b + f"self_attn.{n}_proj.weight"], out[b + f"self_attn.{n}_proj.bias"] = ww, bb
        out[b + "self_attn.out_proj.weight"], out[b + "self_attn.out_proj.bias"] = sd[a + "attn.out_proj.weight"], sd[a + "attn.out_proj.bias"]
        out[b + "layer_norm1.weight"], out[b + "layer_norm1.bias"] = sd[a + "ln_1.weight"], sd[a + "ln_1.bias"]
        out[b + "layer_norm2.weight"], out[b + "layer_norm2.bias"] = sd[a + "ln_2.weight"], sd[a + "ln_2.bias"]
        out[b + "mlp.fc1.weight"], out[b + "mlp.fc1.bias"] = sd[a + "mlp.c_fc.weight"], sd[a + "mlp.c_fc.bias"]
        out[b + "mlp.fc2.weight"], out[b + "mlp.fc2.bias"] = sd[a + "mlp.c_proj.weight"], sd[a + "mlp.c_proj.bias"]
    return out


def main():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    sd = {k: torch.from_numpy(v) for k, v in S.clip_state().items()}
    cfg = CLIPVisionConfig(hidden_size=1024, intermediate_size=4096, num_hidden_layers=24, num_attention_heads=16, image_size=224,
                           patch_size=14, projection_dim=768, hidden_act="quick_gelu", layer_norm_eps=1e-5, attention_dropout=0.0)
    hf = CLIPVisionModelWithProjection(cfg).eval()
    res = hf.load_state_dict(to_hf(sd), strict=False)
    assert not res.unexpected_keys and all("position_ids" in k for k in res.missing_keys), res
    x = torch.from_numpy(clip_input())
    with torch.no_grad():
        x224 = CO.preprocess(x)
        want = hf(pixel_values=x224).image_embeds
        got = CO.encode_image(sd, x224)
    err = float((want - got).abs().max())
    print(f"A8 CLIP ViT-L/14 image embedding: max|HF - oracle| = {err:.2e}  (|embedding| max {float(want.abs().max()):.3f}, std {float(want.std()):.3f})")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "clip_mini.npz"), embed=want.numpy(),
                        x224=x224.numpy()[:, :, ::8, ::8])
    return 0 if err < 1e-4 else 1


if __name__ == "__main__":
    sys.exit(main())
