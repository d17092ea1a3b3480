"""sdf_query forward / gradient throughput for both kernels on 2^20 random points inside a synthetic 96^3 volume."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from o2345 import ops, synthetic as S, _lib as L
from o2345.pipeline import build_networks
dev = torch.device("cuda:0")
tr = build_networks(dev, vol_dim=96, states=S.all_states(0), perturb=0.0)
pack = tr.sdf_network_lod0.sdf_layer.packed()
g = torch.Generator(device="cuda").manual_seed(0)
vol = torch.randn(96, 96, 96, 16, device=dev, generator=g) * 0.3
pts = (torch.rand(1 << 20, 3, device=dev, generator=g) * 1.9 - 0.95)
src = ops.PointSource.explicit(pts)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ref = {}
for prec in (0, 1):
    for name, kw in (("sdf only", {}), ("sdf+feat", {"want_feat": True}), ("sdf+grad", {"want_grad": True})):
        f = lambda: ops.sdf_query(src, vol, pack, precision=prec, **kw)
        out = f(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[2]
        msg = f"precision {prec} {name:9s}: {ms:7.3f} ms  {pts.shape[0] / ms / 1e3:7.1f} M points/s"
        if prec == 0:
            ref[name] = out
        else:
            msg += "   max |sdf - fp32 kernel| = %.2e" % float((out["sdf"] - ref[name]["sdf"]).abs().max())
            if "grad" in out:
                msg += "  grad %.2e" % float((out["grad"] - ref[name]["grad"]).abs().max())
        print(msg)
