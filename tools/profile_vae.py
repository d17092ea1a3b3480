"""One eager VAE decode of 4 latents (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import torch
from o2345.autoencoder import AutoencoderKL
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
vae = AutoencoderKL().cuda().requires_grad_(False)
z = torch.randn(B, 4, 32, 32, device="cuda")
vae.decode(z); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); vae.decode(z); e1.record(); torch.cuda.synchronize()
print("decode(%d): %.3f ms" % (B, e0.elapsed_time(e1)))
