"""Per-shape A/B of the per-tile kernel against the persistent kernel (each tile width), on the big-batch UNet shapes:
20 launches back to back inside one CUDA graph (operands stay L2-resident where they fit, as in the live UNet graph)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from o2345 import _lib as L, ops_a as A
lib = L.load()
# (kind, M, N, K, act, residual)   kind conv: M = B*H*W with H = W given by the level
SHAPES = [("gemm", 65536, 2560, 320, 3, 0), ("gemm", 65536, 320, 1280, 0, 1), ("gemm", 65536, 960, 320, 0, 0), ("gemm", 65536, 320, 320, 0, 1),
          ("conv", 65536, 320, 2880, 0, 1), ("conv", 16384, 640, 5760, 0, 1), ("gemm", 16384, 5120, 640, 3, 0), ("gemm", 16384, 640, 2560, 0, 1),
          ("gemm", 16384, 1920, 640, 0, 0), ("gemm", 16384, 640, 640, 0, 1), ("gemm", 4096, 10240, 1280, 3, 0), ("conv", 4096, 1280, 11520, 0, 1),
          ("gemm", 4096, 1280, 5120, 0, 1), ("gemm", 4096, 3840, 1280, 0, 0), ("conv", 65536, 320, 5760, 0, 0), ("gemm", 65536, 640, 5760, 0, 0),
          ("gemm", 16384, 2560, 320, 3, 0), ("gemm", 16384, 320, 1280, 0, 1), ("gemm", 16384, 320, 320, 0, 1), ("conv", 16384, 320, 2880, 0, 1),
          ("conv", 4096, 640, 5760, 0, 1), ("gemm", 4096, 5120, 640, 3, 0)]
side = torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timed(fn, reps=20):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps): fn()
    g.replay(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) / reps * 1e3
print("%-5s %6s %6s %6s act res | %8s %8s | %8s %8s %8s | %7s" % ("kind", "M", "N", "K", "per-tile", "(bn)", "pers128", "pers160", "pers256", "cuBLAS"))
for kind, M, N, K, act, res in SHAPES:
    gen = torch.Generator(device="cuda").manual_seed(1)
    w = (torch.randn(N, K, device="cuda", generator=gen) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=gen)
    r = torch.randn(M, N, device="cuda", generator=gen).half() if res else None
    if kind == "conv":
        C = K // 9
        HW = {65536: 32, 16384: 16, 4096: 8}[M] if M in (65536, 16384, 4096) else 32
        Bn = M // (HW * HW)
        x = (torch.randn(M, C, device="cuda", generator=gen) * 0.5).half()
        fn = lambda: A.conv3x3(x, Bn, HW, HW, C, w, bias=bias, residual=r)
    else:
        a = (torch.randn(M, K, device="cuda", generator=gen) * 0.5).half()
        fn = lambda: A.gemm(a, w, bias=bias, residual=r, act=act)
    row = []
    lib.o2345_debug_gemm_force(0, 0, 0); lib.o2345_debug_gemm_persist(2, 0)
    base = timed(fn)
    best_bn = ""
    for bn in (128, 160, 256):
        lib.o2345_debug_gemm_force(2, bn, 1); lib.o2345_debug_gemm_persist(2, 0)
        t = timed(fn)
        if t < base * 0.999: base, best_bn = t, str(bn)
        lib.o2345_debug_gemm_persist(1, 0)
        row.append(timed(fn))
    lib.o2345_debug_gemm_force(0, 0, 0); lib.o2345_debug_gemm_persist(0, 0)
    ta = torch.randn(M, K, device="cuda", dtype=torch.float16); to = torch.empty(M, N, device="cuda", dtype=torch.float16)
    lib_us = timed(lambda: torch.matmul(ta, w.t(), out=to))
    print("%-5s %6d %6d %6d %3d %3d | %8.1f %8s | %8.1f %8.1f %8.1f | %7.1f" % (kind, M, N, K, act, res, base, best_bn or "heur", row[0], row[1], row[2], lib_us), flush=True)
    del ta, to
