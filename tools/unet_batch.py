"""UNet iteration time (captured graph, queued calls) against the batch: what batching the reference's independent
sampler calls buys (zero123.generate_views)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import torch
from o2345.unet import UNetModel
net = UNetModel().cuda().requires_grad_(False)
for B in [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]:
    x = torch.randn(B, 8, 32, 32, device="cuda"); t = torch.full((B,), 501, device="cuda"); ctx = torch.randn(B, 1, 768, device="cuda")
    t0 = time.time(); net(x, t, ctx); torch.cuda.synchronize(); cap = time.time() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net(x, t, ctx); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): net(x, t, ctx)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("batch %3d: %.3f ms per iteration = %.3f ms per 8 (capture %.1f s, peak mem %.1f GB)" % (B, ms, ms * 8 / B, cap, torch.cuda.max_memory_allocated() / 2**30), flush=True)
