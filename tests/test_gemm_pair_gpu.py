"""GPU parity of the CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM kernel."""
import pytest
import torch
import torch.nn.functional as F

from test_gemm_gpu import _setup, ref_conv, rel_l2

pytestmark = pytest.mark.gpu


PAIR_CASES = [
    # name, B, H, W, [(C, taps)], N, out_mode, bias, resid, alpha, bn
    ("pair_conv3_w128", 2, 128, 128, [(192, 9)], 192, 0, True, True, 0.70710678, 192),
    ("pair_conv3_w64_n384", 3, 64, 64, [(192, 9)], 384, 0, True, False, 1.0, 192),
    ("pair_conv3_w32_bn256", 3, 32, 32, [(384, 9)], 768, 0, True, False, 1.0, 256),
    ("pair_fused_skip_w16", 5, 16, 16, [(576, 9), (384, 1), (192, 1)], 576, 0, True, False, 0.70710678, 192),
    ("pair_odd_tiles_w8", 5, 8, 8, [(192, 9)], 768, 1, True, False, 1.0, 128),
    ("pair_nin_f32", 2, 32, 32, [(384, 1)], 384, 1, True, False, 1.0, 64),
    ("pair_stride2", 2, 32, 32, [(64, 9)], 128, 0, True, False, 1.0, 128),
    ("pair_final_conv_n15", 2, 128, 128, [(192, 9)], 15, 1, True, False, 1.0, 16),
    ("pair_n40_bn48", 3, 32, 32, [(96, 9)], 40, 0, True, False, 1.0, 48),
    ("pair_ragged_n200_bn64", 3, 16, 16, [(192, 9)], 200, 0, True, True, 1.0, 64),
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_cta_pair_gemm(case):
    """tcgen05 cta_group::2 variant (256-row tiles across two SMs, B tile split between them), incl. fused
    GroupNorm statistics, against torch and against the single-CTA variant (bit-identical accumulation order)."""
    ops = _setup()
    name, B, H, W, segspec, N, out_mode, use_bias, use_resid, alpha, bn = case
    stride = 2 if "stride2" in name else 1
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 2**31)
    segs = [(torch.randn(B, H * stride, W * stride, Cc, device="cuda", generator=g).bfloat16(), taps) for Cc, taps in segspec]
    K = sum(Cc * taps for Cc, taps in segspec)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) if use_bias else None
    resid = torch.randn(B, H, W, N, device="cuda", generator=g).bfloat16() if use_resid else None
    outs, stats = [], []
    for cg in (1, 2):
        out = torch.full((B, H, W, N), float("nan"), device="cuda", dtype=torch.bfloat16 if out_mode == 0 else torch.float32)
        st = torch.zeros(B, N, 2, device="cuda", dtype=torch.int64) if (out_mode == 0 and (H * W) % 32 == 0) else None
        plan = ops.GemmPlan(segs, w, out, out_mode, out_ld=N, bias=bias, resid=resid, resid_ld=N, alpha=alpha, bn=bn,
                            stats=st, stride=stride, cta_group=cg)
        assert plan.cta_group == cg
        plan.launch()
        plan.launch()  # relaunch: barriers / TMEM are re-initialised per launch
        torch.cuda.synchronize()
        outs.append(out)
        stats.append(st)
    assert torch.isfinite(outs[1].float()).all(), f"{name}: non-finite / unwritten outputs"
    if stride == 1:
        ref = ref_conv(segs, w, bias, resid, alpha)
    else:
        a = segs[0][0].float().permute(0, 3, 1, 2)
        ww = w.float().reshape(N, 3, 3, -1).permute(0, 3, 1, 2)
        ref = (F.conv2d(a, ww, stride=2, padding=1) + bias.view(1, -1, 1, 1)) * alpha
    got = outs[1].float().permute(0, 3, 1, 2)
    assert rel_l2(got, ref) < (4e-3 if out_mode == 0 else 2e-5), f"{name}: rel-L2 {rel_l2(got, ref):.3e}"
    assert torch.equal(outs[0], outs[1]), f"{name}: pair variant differs from single-CTA variant"
    if stats[1] is not None:
        assert torch.equal(stats[0], stats[1])
