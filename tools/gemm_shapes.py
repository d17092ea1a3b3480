"""Per-shape device time of every tensor-core GEMM / implicit conv of one UNet forward (CFG batch 8).

Records the C-ABI calls of one eager forward, then replays each distinct shape inside a CUDA graph of
REPS x (L2 flush, call) and subtracts a graph of REPS flushes: device time per call with a cold L2 and no
host launch overhead -- the conditions inside the captured UNet graph.  Buffers are whatever the caching
allocator still holds at those addresses (values are irrelevant for timing).
"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "one-2-3-45_b200")):
    sys.path.insert(0, p)
import torch
from o2345 import _lib as L
from o2345.unet import UNetModel
from o2345.ops import _stream

REPS = 10
net = UNetModel().cuda().requires_grad_(False)
net.use_cuda_graph = False
BATCH = int(os.environ.get("UNET_BATCH", "8"))
x = torch.randn(BATCH, 8, 32, 32, device="cuda"); t = torch.full((BATCH,), 501, device="cuda"); ctx = torch.randn(BATCH, 1, 768, device="cuda")
net(x, t, ctx); torch.cuda.synchronize()

import o2345.ops_a as A
rec = []                                   # (fn, args, kwargs) with the operand tensors kept alive
_gemm, _conv = A.gemm, A.conv3x3
def spy_gemm(*a, **k):
    rec.append(("gemm", _gemm, a, k)); return _gemm(*a, **k)
def spy_conv(*a, **k):
    rec.append(("conv", _conv, a, k)); return _conv(*a, **k)
A.gemm, A.conv3x3 = spy_gemm, spy_conv
keep = net(x, t, ctx); torch.cuda.synchronize()
A.gemm, A.conv3x3 = _gemm, _conv

def key(kind, a, k):
    rb, res, act = int(k.get("rowbias") is not None), int(k.get("residual") is not None), int(k.get("act", 0))
    if kind == "gemm":
        return ("gemm", a[0].shape[0], a[1].shape[0], a[0].shape[1], 1, rb, res, act)
    B, H, W, C = a[1:5]
    return ("conv", B * H * W, a[5].shape[0], 9 * C, 1, rb, res, act)

groups = collections.OrderedDict()
for kind, fn, a, k in rec:
    groups.setdefault(key(kind, a, k), []).append((fn, a, k))

flush = torch.empty(64 << 20, dtype=torch.float32, device="cuda")   # 256 MB > L2
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
side = torch.cuda.Stream()

def graph_ms(fn):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=side):
            for _ in range(REPS):
                flush.zero_()
                fn()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[1] / REPS

base = graph_ms(lambda: None)
rows = []
for k, calls in groups.items():
    fn, a, kw = calls[0]
    ms = max(graph_ms(lambda: fn(*a, **kw)) - base, 1e-4)
    kind, M, N, K, batch = k[:5]
    lib_ms = float("nan")
    if batch == 1:                          # diagnostic ceiling: the vendor library on the same shape, same protocol
        ta = torch.randn(M, K, device="cuda", dtype=torch.float16); tb = torch.randn(N, K, device="cuda", dtype=torch.float16)
        to = torch.empty(M, N, device="cuda", dtype=torch.float16)
        lib_ms = max(graph_ms(lambda: torch.matmul(ta, tb.t(), out=to)) - base, 1e-4)
        del ta, tb, to
    rows.append((ms * len(calls), len(calls), ms, 2.0 * M * N * K * batch, k, lib_ms))
tot = sum(r[0] for r in rows); totfl = sum(r[3] * r[1] for r in rows)
libtot = sum(r[5] * r[1] for r in rows if r[5] == r[5])
print(f"flush baseline {base*1e3:.1f} us; library ceiling on the non-batched shapes: {libtot:.3f} ms")
print(f"total {tot:.3f} ms for {totfl/1e12:.3f} TFLOP -> {totfl/tot/1e9:.1f} TFLOP/s   ({len(rec)} calls, {len(rows)} shapes)")
print(f"{'tot ms':>8s} {'n':>3s} {'us':>8s} {'TF/s':>7s}  kind      M      N      K  batch rowb res act   (cuBLAS same shape)")
for r in sorted(rows, reverse=True):
    t_, n, ms, fl, k, lib_ms = r
    print(f"{t_:8.3f} {n:3d} {ms*1e3:8.1f} {fl/ms/1e9:7.1f}  {k[0]:5s} {k[1]:6d} {k[2]:6d} {k[3]:6d} {k[4]:6d} {k[5]:4d} {k[6]:3d} {k[7]:3d}   lib {lib_ms*1e3:7.1f} us")
