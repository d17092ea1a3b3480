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

REPS = 10
net = UNetModel().cuda()
net.use_cuda_graph = False
x = torch.randn(8, 8, 32, 32, device="cuda"); t = torch.full((8,), 501, device="cuda"); ctx = torch.randn(8, 1, 768, device="cuda")
net(x, t, ctx); torch.cuda.synchronize()

rec = []
orig = L.call
def spy(name, *args):
    if name in ("o2345_gemm_f16", "o2345_conv3x3_f16"):
        rec.append((name, args))
    return orig(name, *args)
L.call = spy
keep = net(x, t, ctx); torch.cuda.synchronize()
L.call = orig

def key(name, args):
    if name == "o2345_gemm_f16":
        M, N, K = args[3:6]; nh, nb = args[9], args[10]; ep = args[17]._obj
        return ("gemm", M, N, K, nh * nb if nh else 1, int(bool(ep.rowbias)), int(bool(ep.residual)), ep.act)
    B, H, W, C = args[1:5]; N = args[6]; ep = args[9]._obj
    return ("conv", B * H * W, N, 9 * C, 1, int(bool(ep.rowbias)), int(bool(ep.residual)), ep.act)

groups = collections.OrderedDict()
for name, args in rec:
    groups.setdefault(key(name, args), []).append((name, args))

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
    name, args = calls[0]
    ms = max(graph_ms(lambda: orig(name, *args)) - base, 1e-4)
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
