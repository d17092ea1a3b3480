"""Quick device timings of the path-A building blocks (UNet step at the CFG batch of 8, VAE decode of 4 latents)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import torch
from o2345.zero123 import build_zero123
from o2345 import _lib

m = build_zero123("cuda")
unet, vae = m.model.diffusion_model, m.first_stage_model
x = torch.randn(8, 8, 32, 32, device="cuda"); t = torch.full((8,), 501, device="cuda"); ctx = torch.randn(8, 1, 768, device="cuda")
z = torch.randn(4, 4, 32, 32, device="cuda")
def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - w0) / n * 1e3
_lib.reset_launches(); unet(x, t, ctx); print("kernels per UNet forward:", _lib.launches())
print("UNet batch 8: device %.2f ms, wall %.2f ms" % timeit(lambda: unet(x, t, ctx), 20))
print("VAE decode batch 4: device %.2f ms, wall %.2f ms" % timeit(lambda: vae.decode(z), 5))
img = torch.rand(1, 3, 256, 256, device="cuda") * 2 - 1
print("VAE encode batch 1: device %.2f ms, wall %.2f ms" % timeit(lambda: vae.encode(img), 5))
