"""One eager UNet forward at batch argv[1] (default 8 = 4 views x CFG) for ncu launch lists."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import torch
from o2345.unet import UNetModel
net = UNetModel().cuda().requires_grad_(False)           # default torch init is fine for timing
net.use_cuda_graph = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
x = torch.randn(B, 8, 32, 32, device="cuda"); t = torch.full((B,), 501, device="cuda"); ctx = torch.randn(B, 1, 768, device="cuda")
net(x, t, ctx); torch.cuda.synchronize()
torch.cuda.nvtx.range_push("unet")
net(x, t, ctx); torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
