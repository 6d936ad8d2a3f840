"""GPU parity of the fused tcgen05 attention kernel against torch fp32 on the same bf16 inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def make_plan(ops, qk, vT, out, heads, scale, v_rows):
    """v_rows: V as rows behind q|k in ONE (B, N, 3C) tensor (the fused q|k|v projection layout, MN-major operand in the
    kernel) instead of the separate V^T (B, C, N) tensor."""
    if not v_rows:
        return ops.AttnPlan(qk, vT, out, heads, scale), qk
    B, N, C2 = qk.shape
    qkv = torch.cat([qk, vT.transpose(1, 2)], -1).contiguous()
    return ops.AttnPlan(qkv, None, out, heads, scale, v=qkv[:, :, C2:]), qkv


# the last four: head dims above 384 (unet.py 'deeper' mid block is one head of 768): O columns split over CTAs, Q streamed
@pytest.mark.parametrize("v_rows", [False, True])
@pytest.mark.parametrize("B,N,C,heads", [(2, 1024, 384, 2), (3, 256, 576, 3), (1, 128, 64, 1), (2, 256, 256, 1),
                                          (1, 4096, 384, 1), (2, 128, 768, 4), (5, 1024, 128, 2), (5, 64, 768, 4), (3, 192, 128, 2),
                                          (2, 64, 768, 1), (3, 256, 768, 1), (1, 1024, 1024, 2), (2, 192, 448, 1)])
def test_fused_attention(B, N, C, heads, v_rows):
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
    plan, _keep = make_plan(ops, qk, vT, out, heads, scale, v_rows)
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


@pytest.mark.parametrize("v_rows", [False, True])
@pytest.mark.parametrize("B,N,C,heads,mode", [(2, 1024, 384, 2, "ramp"), (1, 512, 256, 1, "ramp"), (2, 256, 192, 1, "late_spike"),
                                               (1, 1024, 128, 2, "early_spike"), (1, 512, 768, 1, "ramp")])
def test_fused_attention_online_rescale(B, N, C, heads, mode, v_rows):
    """The one-pass online softmax raises a row's reference lazily (only when a key tile exceeds it by more than 2^8) and
    then rescales the O accumulator in TMEM.  Scores that grow along the key axis force that path in every tile
    ("ramp"), once near the end ("late_spike") or never after the first tile ("early_spike")."""
    torch.backends.cuda.matmul.allow_tf32 = False
    from evcdiff import ops
    g = torch.Generator(device="cuda").manual_seed(N + C)
    d = C // heads
    scale = d ** -0.5
    q = torch.randn(B, N, C, device="cuda", generator=g)
    k = torch.randn(B, N, C, device="cuda", generator=g) * 0.3
    # add a component along each query's own direction so that q.k grows with the key index
    qn = q.reshape(B, N, heads, d)
    u = qn.mean(dim=1, keepdim=True)
    u = u / u.norm(dim=-1, keepdim=True)  # (B,1,heads,d): one direction per head
    qn = qn + 6.0 * u  # every query has a positive component along u
    idx = torch.arange(N, device="cuda", dtype=torch.float32)
    if mode == "ramp":
        amp = 40.0 * max(1.0, (d / 256) ** 0.5) * idx / N  # scores rise by ~ 6 * 40 * scale * log2(e) over the key axis
    elif mode == "late_spike":
        amp = torch.where(idx >= N - 40, 30.0, 0.0)
    else:
        amp = torch.where(idx < 8, 30.0, 0.0)
    kn = k.reshape(B, N, heads, d) + amp.view(1, N, 1, 1) * u
    qk = torch.cat([qn.reshape(B, N, C), kn.reshape(B, N, C)], -1).bfloat16()
    vT = torch.randn(B, C, N, device="cuda", generator=g).bfloat16()
    out = torch.full((B, N, C), float("nan"), device="cuda", dtype=torch.bfloat16)
    plan, _keep = make_plan(ops, qk, vT, out, heads, scale, v_rows)
    plan.launch()
    torch.cuda.synchronize()
    qf = qk[:, :, :C].float().reshape(B, N, heads, d).permute(0, 2, 1, 3)
    kf = qk[:, :, C:].float().reshape(B, N, heads, d).permute(0, 2, 1, 3)
    v = vT.float().reshape(B, heads, d, N).permute(0, 1, 3, 2)
    s = qf @ kf.transpose(-1, -2) * scale
    if mode == "ramp":  # the case really spans many rescale thresholds
        span = (s.max(-1).values - s[..., :64].max(-1).values) * 1.4427
        assert float(span.median()) > 16.0, float(span.median())  # at least two raises of the reference per row
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, N, C)
    assert torch.isfinite(out.float()).all()
    err = rel_l2(out.float(), ref)
    assert err < 8e-3, err
