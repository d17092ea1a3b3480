"""GPU: the tcgen05 GEMM against torch.matmul in fp32 on the same fp16 inputs (fp32 accumulate on both sides)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def ref(a, b, bias=None, residual=None, act=0, alpha=1.0):
    y = alpha * (a.float() @ b.float().t())
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.silu(y)
    if act == 2:
        y = torch.nn.functional.gelu(y)
    if residual is not None:
        y = y + residual.float()
    return y


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 256), (8192, 320, 2880), (100, 72, 40), (512, 1280, 11520),
                                   (64, 2560, 320), (1, 1280, 320), (333, 200, 136)])
def test_gemm_matches_matmul(M, N, K):
    from o2345 import ops_a
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    b = (torch.randn(N, K, device="cuda", generator=g) * 0.5).half()
    out = ops_a.gemm(a, b, out_dtype=torch.float32)
    torch.cuda.synchronize()
    want = ref(a, b)
    err = (out - want).abs().max().item()
    assert err <= 2e-3 * (K ** 0.5), (err, M, N, K)          # fp32 accumulation: only summation order differs


def test_gemm_epilogues_and_fp16_out():
    from o2345 import ops_a
    g = torch.Generator(device="cuda").manual_seed(1)
    a = (torch.randn(300, 192, device="cuda", generator=g) * 0.3).half()
    b = (torch.randn(160, 192, device="cuda", generator=g) * 0.3).half()
    bias = torch.randn(160, device="cuda", generator=g)
    res = torch.randn(300, 160, device="cuda", generator=g).half()
    for act in (0, 1, 2):
        out = ops_a.gemm(a, b, bias=bias, residual=res, act=act, alpha=0.5)
        want = ref(a, b, bias, res, act, 0.5)
        assert out.dtype == torch.float16
        assert (out.float() - want).abs().max().item() < 2e-2
    # strided A (a column slice of a wider matrix), as the attention q/k/v views use
    wide = (torch.randn(256, 3 * 64, device="cuda", generator=g) * 0.3).half()
    q = wide[:, 64:128]
    out = ops_a.gemm(q, b[:, :64].contiguous(), out_dtype=torch.float32)
    assert (out - ref(q, b[:, :64])).abs().max().item() < 1e-2


def test_batched_gemm_heads_inside_a_token_tensor():
    """q, k live as column blocks of a [B, N, 3*H*d] tensor (fused qkv projection); scores per (b, h)."""
    from o2345 import ops_a
    g = torch.Generator(device="cuda").manual_seed(2)
    B, N, H, d = 2, 256, 8, 40
    Cc = H * d
    qkv = (torch.randn(B, N, 3 * Cc, device="cuda", generator=g) * 0.5).half()
    q, k = qkv[:, :, :Cc], qkv[:, :, Cc:2 * Cc]
    out = torch.empty(B * H, N, N, dtype=torch.float16, device="cuda")
    ops_a.bgemm(q, k, out, H, B, (d, N * 3 * Cc), (d, N * 3 * Cc), (N * N, H * N * N), N, N, d, 3 * Cc, 3 * Cc, N, alpha=d ** -0.5)
    want = torch.einsum("bihd,bjhd->bhij", q.float().view(B, N, H, d), k.float().view(B, N, H, d)) * d ** -0.5
    assert (out.float().view(B, H, N, N) - want).abs().max().item() < 2e-2
    # P @ V with V^T [B, C, N] and the result written back into [B, N, C]
    p = torch.softmax(want, -1).half().contiguous().view(B * H, N, N)
    v = qkv[:, :, 2 * Cc:]
    vt = ops_a.transpose_tokens(v.contiguous(), B, N, Cc)
    o = torch.zeros(B, N, Cc, dtype=torch.float16, device="cuda")
    ops_a.bgemm(p, vt, o, H, B, (N * N, H * N * N), (d * N, Cc * N), (d, N * Cc), N, d, N, N, N, Cc)
    want_o = torch.einsum("bhij,bjhd->bihd", p.float().view(B, H, N, N), v.float().view(B, N, H, d)).reshape(B, N, Cc)
    assert (o.float() - want_o).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,H,W,C,N", [(8, 32, 32, 320, 320), (8, 16, 16, 640, 640), (8, 8, 8, 1280, 1280), (8, 4, 4, 1280, 1280),
                                       (2, 64, 64, 512, 512), (1, 256, 256, 128, 128), (3, 8, 8, 72, 40), (1, 128, 128, 256, 3)])
def test_implicit_conv3x3_matches_conv2d(B, H, W, C, N):
    """The TMA shifted-window convolution against F.conv2d (fp32) on the same fp16 operands."""
    from o2345 import ops_a
    g = torch.Generator(device="cuda").manual_seed(H + C)
    x = (torch.randn(B, H, W, C, device="cuda", generator=g) * 0.5).half()
    w = (torch.randn(N, C, 3, 3, device="cuda", generator=g) * (2.0 / (9 * C)) ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g)
    wk = w.permute(0, 2, 3, 1).reshape(N, 9 * C).contiguous()
    out = ops_a.conv3x3(x.view(-1, C), B, H, W, C, wk, bias=bias, out_dtype=torch.float32)
    want = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).float(), w.float(), bias, padding=1).permute(0, 2, 3, 1).reshape(-1, N)
    err = (out - want).abs().max().item()
    assert err < 5e-3, err
