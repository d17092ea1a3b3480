"""Path A (Zero123 DDIM / UNet / VAE) tensor-level wrappers over the C-ABI: fp16 activations, fp32 norms."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .ops import _p, _stream

_f16, _f32 = torch.float16, torch.float32


def gemm(a, b, bias=None, residual=None, act=0, alpha=1.0, out_dtype=_f16, out=None):
    """out[M,N] = act(alpha * a[M,K] @ b[N,K]^T + bias) + residual.  a, b fp16 with contiguous K; rows may be strided."""
    assert a.dtype == _f16 and b.dtype == _f16 and a.stride(-1) == 1 and b.stride(-1) == 1
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    assert out.stride(-1) == 1
    L.call("o2345_gemm_f16", C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), M, N, K,
           a.stride(0), b.stride(0), out.stride(0), 0, 0, 0, 0, _p(bias, _f32),
           None if residual is None else C.c_void_p(residual.data_ptr()), int(act), float(alpha),
           int(out.dtype == _f32), _stream())
    return out


def bgemm(a, b, alpha=1.0, out_dtype=_f16):
    """Batched: out[B,M,N] = alpha * a[B,M,K] @ b[B,N,K]^T (attention scores / PV)."""
    assert a.dtype == _f16 and b.dtype == _f16 and a.stride(-1) == 1 and b.stride(-1) == 1
    Bn, M, K = a.shape
    N = b.shape[1]
    out = torch.empty(Bn, M, N, dtype=out_dtype, device=a.device)
    L.call("o2345_gemm_f16", C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(out.data_ptr()), M, N, K,
           a.stride(1), b.stride(1), out.stride(1), Bn, a.stride(0), b.stride(0), out.stride(0), None, None, 0,
           float(alpha), int(out_dtype == _f32), _stream())
    return out
