"""Runs one GEMM shape a few times (ncu target): python tools/gemm_one.py M N K [residual]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import torch
from o2345 import ops_a as A
M, N, K = (int(v) for v in sys.argv[1:4])
res = len(sys.argv) > 4 and sys.argv[4] != "0"
a = torch.randn(M, K, device="cuda").half(); b = torch.randn(N, K, device="cuda").half()
bias = torch.randn(N, device="cuda"); r = torch.randn(M, N, device="cuda").half() if res else None
for _ in range(5):
    A.gemm(a, b, bias=bias, residual=r)
torch.cuda.synchronize()
