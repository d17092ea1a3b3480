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


def test_batched_gemm():
    from o2345 import ops_a
    g = torch.Generator(device="cuda").manual_seed(2)
    a = (torch.randn(16, 256, 40, device="cuda", generator=g) * 0.5).half()
    b = (torch.randn(16, 256, 40, device="cuda", generator=g) * 0.5).half()
    out = ops_a.bgemm(a, b, alpha=40 ** -0.5, out_dtype=torch.float32)
    want = torch.einsum("bik,bjk->bij", a.float(), b.float()) * 40 ** -0.5
    assert (out - want).abs().max().item() < 5e-3
