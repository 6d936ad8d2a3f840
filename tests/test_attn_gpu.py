"""GPU parity of the fused tcgen05 attention kernel against torch fp32 on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("B,N,C,heads", [(2, 1024, 384, 2), (3, 256, 576, 3), (1, 128, 64, 1), (2, 256, 256, 1),
                                          (1, 4096, 384, 1), (2, 128, 768, 4), (5, 1024, 128, 2), (5, 64, 768, 4), (3, 192, 128, 2)])
def test_fused_attention(B, N, C, heads):
    torch.backends.cuda.matmul.allow_tf32 = False
    from evcdiff import ops
    assert ops.attn_supported(N, C, heads)
    g = torch.Generator(device="cuda").manual_seed(N + C + heads)
    d = C // heads
    qk = torch.randn(B, N, 2 * C, device="cuda", generator=g).bfloat16()
    qk[:, :, :C] *= 2.0  # sharper softmax: exercises the max subtraction
    vT = torch.randn(B, C, N, device="cuda", generator=g).bfloat16()
    out = torch.full((B, N, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = d ** -0.5
    plan = ops.AttnPlan(qk, vT, out, heads, scale)
    plan.launch()
    plan.launch()
    torch.cuda.synchronize()
    q = qk[:, :, :C].float().reshape(B, N, heads, d).permute(0, 2, 1, 3)
    k = qk[:, :, C:].float().reshape(B, N, heads, d).permute(0, 2, 1, 3)
    v = vT.float().reshape(B, heads, d, N).permute(0, 1, 3, 2)
    ref = torch.softmax(q @ k.transpose(-1, -2) * scale, -1) @ v  # (B, heads, N, d)
    ref = ref.permute(0, 2, 1, 3).reshape(B, N, C)
    assert torch.isfinite(out.float()).all()
    err = rel_l2(out.float(), ref)
    assert err < 8e-3, err  # P is rounded to bf16 before the PV product (like the unfused path), output to bf16
