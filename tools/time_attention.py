"""Self-attention kernel at the UNet's shapes for the batched sampler calls (20 launches back to back in a graph)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from o2345 import ops_a as A
side = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for B in (16, 64):
    for N, H, d in ((1024, 8, 40), (256, 8, 80), (64, 8, 160)):
        C = H * d
        qkv = (torch.randn(B * N, 3 * C, device="cuda") * 0.5).half()
        fn = lambda: A.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, N, H, d)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fn(); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                for _ in range(20): fn()
        g.replay(); torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / 20 * 1e3)
        us = float(np.median(ts))
        print("B %3d N %5d heads %d d %3d: %8.1f us  %6.1f TFLOP/s (4 N^2 d per head)" % (B, N, H, d, us, 4.0 * N * N * d * H * B / us / 1e6), flush=True)
