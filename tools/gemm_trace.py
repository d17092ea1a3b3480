"""Phase timeline (SM cycles) of CTA (0,0,0) of the pair GEMM for a few UNet shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import ctypes as C
import torch
from o2345 import _lib as L, ops_a as A
buf = torch.zeros(16, dtype=torch.int64, device="cuda")
names = ["entry", "prologue", "tma0", "tmaN", "landed0", "mmaN", "acc", "epi", "exit"]
flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
for M, N, K, res in [(8192, 320, 320, 0), (8192, 320, 320, 1), (2048, 640, 640, 0), (8192, 2560, 320, 0), (2048, 640, 5760, 0), (512, 1280, 1280, 0)]:
    a = torch.randn(M, K, device="cuda").half(); b = torch.randn(N, K, device="cuda").half()
    bias = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda").half() if res else None
    for cold in (0, 1):
        A.gemm(a, b, bias=bias, residual=r)
        if cold: flush.zero_()
        torch.cuda.synchronize()
        L.load().o2345_debug_gemm_trace(C.c_void_p(buf.data_ptr()))
        A.gemm(a, b, bias=bias, residual=r)
        torch.cuda.synchronize()
        L.load().o2345_debug_gemm_trace(None)
        t = buf.tolist()
        print(f"M={M} N={N} K={K} res={res} cold={cold}: " + "  ".join(f"{n}={t[i]-t[0]}" for i, n in enumerate(names)))
