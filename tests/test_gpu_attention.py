"""GPU: fused self-attention kernel against a torch fp32 reference of the same op on the same fp16 inputs
(tolerance 3e-3 abs on outputs of O(1): fp16 P and V operands, fp32 accumulation)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,H,d", [(2, 1024, 8, 40), (8, 256, 8, 80), (3, 64, 8, 160), (2, 16, 8, 160), (1, 100, 2, 40), (1, 640, 4, 80),
                                     (1, 1000, 2, 160), (2, 257, 16, 64), (1, 577, 3, 64)])
def test_attention_matches_reference(B, N, H, d):
    from o2345 import ops_a
    g = torch.Generator(device="cuda").manual_seed(N + d)
    C = H * d
    qkv = (torch.randn(B * N, 3 * C, device="cuda", generator=g) * 0.8).half()
    out = ops_a.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B, N, H, d)
    torch.cuda.synchronize()
    q, k, v = (qkv[:, i * C:(i + 1) * C].float().view(B, N, H, d).permute(0, 2, 1, 3) for i in range(3))
    want = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, -1) @ v
    want = want.permute(0, 2, 1, 3).reshape(B * N, C)
    err = (out.float() - want).abs().max().item()
    assert err < 3e-3, err
