"""CPU restatement of the Zero123 conditioning encoder -- TEST INFRASTRUCTURE ONLY (tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline may import it; the product path never does).

Follows reference ldm/modules/encoders/modules.py:362-379 (`FrozenCLIPImageEmbedder.preprocess / forward`) and, for
`model.encode_image`, the published OpenAI CLIP vision tower (github.com/openai/CLIP @ clip/model.py: `VisionTransformer`
:206-240, `ResidualAttentionBlock` :171-193, `QuickGELU` :166-168, `LayerNorm` computing in fp32 :157-163).  The `clip`
and `kornia` packages are requirements.txt dependencies that are NOT under /root/reference (and not installed here):
parity is pinned on Hugging Face transformers' independent implementation of the same architecture
(oracle/pin_clip_against_hf.py) instead of on the reference's own run -- "parity unpinned" against the reference
itself for this row.  kornia.geometry.resize(..., 'bicubic', align_corners=True, antialias=False) is
torch.nn.functional.interpolate with the same arguments (kornia/geometry/transform/affwarp.py `resize`).
"""
import torch
import torch.nn.functional as F

MEAN = torch.tensor([0.48145466, 0.4578275, 0.40821073])
STD = torch.tensor([0.26862954, 0.26130258, 0.27577711])


def preprocess(x):
    """[-1, 1] images [B,3,H,W] -> normalised 224 x 224 (modules.py:362-370)."""
    x = F.interpolate(x, size=(224, 224), mode="bicubic", align_corners=True, antialias=False)
    x = (x + 1.0) / 2.0
    return (x - MEAN[None, :, None, None]) / STD[None, :, None, None]


def encode_image(sd, x, heads=16, prefix="model.visual."):
    """OpenAI CLIP VisionTransformer.forward (clip/model.py:223-240) with the state dict `sd` (keys `model.visual.*`)."""
    g = lambda k: sd[prefix + k]
    x = F.conv2d(x, g("conv1.weight"), stride=g("conv1.weight").shape[-1])              # [B, width, 16, 16]
    B, d = x.shape[0], x.shape[1]
    x = x.reshape(B, d, -1).permute(0, 2, 1)                                             # [B, 256, width]
    x = torch.cat([g("class_embedding")[None, None].expand(B, 1, d), x], 1) + g("positional_embedding")[None]
    x = F.layer_norm(x, (d,), g("ln_pre.weight"), g("ln_pre.bias"))
    layers = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith(prefix + "transformer.resblocks."))
    hd = d // heads
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        y = F.layer_norm(x, (d,), g(p + "ln_1.weight"), g(p + "ln_1.bias"))
        qkv = y @ g(p + "attn.in_proj_weight").t() + g(p + "attn.in_proj_bias")
        q, k, v = (t.reshape(B, -1, heads, hd).transpose(1, 2) for t in qkv.chunk(3, -1))   # [B, H, N, hd]
        a = torch.softmax((q @ k.transpose(-1, -2)) * hd ** -0.5, -1) @ v
        a = a.transpose(1, 2).reshape(B, -1, d)
        x = x + a @ g(p + "attn.out_proj.weight").t() + g(p + "attn.out_proj.bias")
        y = F.layer_norm(x, (d,), g(p + "ln_2.weight"), g(p + "ln_2.bias"))
        y = y @ g(p + "mlp.c_fc.weight").t() + g(p + "mlp.c_fc.bias")
        y = y * torch.sigmoid(1.702 * y)                                                   # QuickGELU
        x = x + y @ g(p + "mlp.c_proj.weight").t() + g(p + "mlp.c_proj.bias")
    x = F.layer_norm(x[:, 0], (d,), g("ln_post.weight"), g("ln_post.bias"))
    return x @ g("proj")


def embed(sd, x):
    """FrozenCLIPImageEmbedder.forward: [-1, 1] images -> [B, 768] (modules.py:372-379)."""
    return encode_image(sd, preprocess(x)).float()
