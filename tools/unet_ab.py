"""Device time of the captured UNet graph (CFG batch 8) under the knobs given in the environment:
O2345_PDL=0, O2345_GEMM_FORCE=ctas,bn,splits, O2345_FUSE_GN=0.  Prints one line."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import torch
from o2345 import _lib
from o2345.unet import UNetModel
net = UNetModel().cuda().requires_grad_(False)
if "O2345_FUSE_GN" in os.environ:
    net.fuse_gn_stats = os.environ["O2345_FUSE_GN"] != "0"
if "O2345_GN_ONE" in os.environ:
    net.gn_one_kernel = os.environ["O2345_GN_ONE"] != "0"
x = torch.randn(8, 8, 32, 32, device="cuda"); t = torch.full((8,), 501, device="cuda"); ctx = torch.randn(8, 1, 768, device="cuda")
for _ in range(3):
    net(x, t, ctx)
torch.cuda.synchronize()
graph = net._graphs[next(iter(net._graphs))][0]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for _ in range(5):
    a.record()
    for _ in range(10):
        graph.replay()
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 10)
_lib.reset_launches(); net(x, t, ctx)
print("UNet graph %.3f ms (min %.3f)  kernels %d  env %s" % (sorted(ts)[2], min(ts), _lib.launches(),
      {k: v for k, v in os.environ.items() if k.startswith("O2345_")}))
