"""GroupNorm (+SiLU) at the UNet's shapes for the sampler batches: the one-kernel cluster version against the two-kernel route
(statistics kernel + apply), 20 launches back to back in one CUDA graph over 4 rotating tensors (L2-warm for the small ones,
as in the live UNet graph)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from o2345 import ops_a as A
side = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(fn, reps=20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(0); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for i in range(reps): fn(i)
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) / reps * 1e3
from o2345 import _lib
lib = _lib.load()
CLS = (1, 2, 4, 8, 16)
print("%4s %6s %6s | %s | %9s" % ("B", "HW", "C", " ".join("cl=%-5d" % c for c in CLS), "two (us)"))
tot = {}
for B in (8, 16, 64):
    for HW, C, n in [(1024, 320, 9), (1024, 640, 2), (1024, 960, 1), (256, 320, 1), (256, 640, 8), (256, 960, 1), (256, 1280, 2), (256, 1920, 1),
                     (64, 640, 1), (64, 1280, 8), (64, 1920, 1), (64, 2560, 2), (16, 1280, 7), (16, 2560, 3)]:
        xs = [torch.randn(B * HW, C, device="cuda").half() for _ in range(4)]
        g, b = torch.randn(C, device="cuda"), torch.randn(C, device="cuda")
        ones = []
        for cl in CLS:
            lib.o2345_debug_groupnorm_cluster(cl)
            ones.append(timed(lambda i: A.groupnorm_apply(xs[i % 4], B, HW, C, 32, 1e-5, g, b, 1)))
        lib.o2345_debug_groupnorm_cluster(0)
        one = min(ones)
        def two(i):
            sc = A.groupnorm_stats(xs[i % 4], B, HW, C, 32, 1e-5, g, b)
            return A.norm_act_im2col(xs[i % 4], B, int(HW ** 0.5), int(HW ** 0.5), C, ksize=1, gn=sc, act=True)
        t2 = timed(two)
        tot[B] = tot.get(B, np.zeros(2)) + n * np.array([one, t2])
        print("%4d %6d %6d | %s | %9.1f" % (B, HW, C, " ".join("%-8.1f" % o for o in ones), t2), flush=True)
        del xs
for B, t in tot.items():
    print("batch %d: all GroupNorms of one UNet pass: one-kernel (best cluster size per shape) %.3f ms, two-kernel %.3f ms" % (B, t[0] / 1e3, t[1] / 1e3))
