"""GPU parity of the tcgen05 implicit-GEMM kernel (evc_gemm_*) against plain torch fp32 on the same bf16 inputs.

Tolerances: fp32 outputs rel-L2 <= 2e-5 (fp32 accumulation, different summation order);
bf16 outputs rel-L2 <= 4e-3 (one bf16 rounding of the result, 2^-9 relative per element).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _setup():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from evcdiff import ops
    return ops


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def ref_conv(segs, w, bias, resid, alpha):
    """segs: list of ((B,H,W,C) bf16, taps); w: (N, K) bf16 with K = seg-major, tap-major, channel-minor."""
    out = None
    k = 0
    for a, taps in segs:
        Cc = a.shape[3]
        x = a.float().permute(0, 3, 1, 2)
        wk = w[:, k:k + taps * Cc].float()
        if taps == 9:
            ww = wk.reshape(-1, 3, 3, Cc).permute(0, 3, 1, 2)
            y = F.conv2d(x, ww, padding=1)
        else:
            y = F.conv2d(x, wk.reshape(-1, Cc, 1, 1))
        out = y if out is None else out + y
        k += taps * Cc
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    if resid is not None:
        out = out + resid.float().permute(0, 3, 1, 2)
    return out * alpha  # NCHW fp32


CASES = [
    # name, B, H, W, [(C, taps)], N, out_mode, bias, resid, alpha
    ("gemm_k64", 1, 1, 128, [(64, 1)], 64, 1, False, False, 1.0),
    ("gemm_k256_n192", 2, 1, 256, [(256, 1)], 192, 1, True, False, 1.0),
    ("conv3_w128", 1, 128, 128, [(64, 9)], 192, 0, True, False, 1.0),
    ("conv3_w64", 2, 64, 64, [(192, 9)], 192, 0, True, True, 0.70710678),
    ("conv3_w32_n384", 3, 32, 32, [(192, 9)], 384, 0, True, False, 1.0),
    ("conv3_w16_n576", 3, 16, 16, [(384, 9)], 576, 0, True, False, 1.0),
    ("conv3_w8_n768", 5, 8, 8, [(576, 9)], 768, 0, True, True, 0.70710678),
    ("conv3_w4", 9, 4, 4, [(128, 9)], 256, 1, True, False, 1.0),
    ("conv3_w2", 33, 2, 2, [(64, 9)], 128, 1, True, False, 1.0),
    ("res_fused_skip", 2, 32, 32, [(384, 9), (192, 1), (192, 1)], 384, 0, True, False, 0.70710678),
    ("conv_out15_nchw", 2, 128, 128, [(192, 9)], 15, 3, True, False, 1.0),
    ("conv_in_pad64", 2, 128, 128, [(64, 9)], 192, 0, True, False, 1.0),
    ("nin_T", 2, 32, 32, [(384, 1)], 384, 2, True, False, 1.0),
    ("many_tiles", 8, 64, 64, [(192, 9)], 192, 0, True, False, 1.0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_gemm(case):
    ops = _setup()
    name, B, H, W, segspec, N, out_mode, use_bias, use_resid, alpha = case
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 2**31)
    segs = [(torch.randn(B, H, W, Cc, device="cuda", generator=g).bfloat16(), taps) for Cc, taps in segspec]
    K = sum(Cc * taps for Cc, taps in segspec)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g) if use_bias else None
    resid = torch.randn(B, H, W, N, device="cuda", generator=g).bfloat16() if use_resid else None
    ref = ref_conv(segs, w, bias, resid, alpha)
    if out_mode in (0, 1):
        out = torch.full((B, H, W, N), float("nan"), device="cuda",
                         dtype=torch.bfloat16 if out_mode == 0 else torch.float32)
        plan = ops.GemmPlan(segs, w, out, out_mode, out_ld=N, bias=bias, resid=resid, resid_ld=N, alpha=alpha)
    else:
        out = torch.full((B, N, H, W), float("nan"), device="cuda",
                         dtype=torch.bfloat16 if out_mode == 2 else torch.float32)
        plan = ops.GemmPlan(segs, w, out, out_mode, out_ld=H * W, out_bs=N * H * W, bias=bias, resid=resid,
                            resid_ld=N, alpha=alpha)
    plan.launch()
    torch.cuda.synchronize()
    got = out.float().permute(0, 3, 1, 2) if out_mode in (0, 1) else out.float()
    assert torch.isfinite(got).all(), f"{name}: non-finite / unwritten outputs"
    err = rel_l2(got, ref)
    tol = 4e-3 if out_mode in (0, 2) else 2e-5
    assert err < tol, f"{name}: rel-L2 {err:.3e} > {tol}"


def test_batched_b_operand():
    """Per-sample B operand (attention QK^T and PV shapes): out[b] = A[b] @ Wb[b]^T."""
    ops = _setup()
    g = torch.Generator(device="cuda").manual_seed(7)
    for (B, M, K, N) in [(3, 1024, 192, 1024), (4, 256, 192, 256), (5, 64, 192, 64), (3, 1024, 1024, 192)]:
        a = torch.randn(B, 1, M, K, device="cuda", generator=g).bfloat16()
        wb = (torch.randn(B, N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
        out = torch.full((B, M, N), float("nan"), device="cuda", dtype=torch.float32)
        plan = ops.GemmPlan([(a, 1)], wb, out, 1, out_ld=N, alpha=0.5)
        plan.launch()
        torch.cuda.synchronize()
        ref = 0.5 * torch.bmm(a[:, 0].float(), wb.float().transpose(1, 2))
        assert torch.isfinite(out).all()
        assert rel_l2(out, ref) < 2e-5, (B, M, K, N, rel_l2(out, ref))


def test_strided_views_and_repeat_launch():
    """A/W given as strided views (head slices of a fused QK buffer); plan relaunch is idempotent."""
    ops = _setup()
    g = torch.Generator(device="cuda").manual_seed(11)
    B, Npix, Cc = 2, 256, 384
    qk = torch.randn(B, Npix, 2 * Cc, device="cuda", generator=g).bfloat16()
    h = 1
    q = qk[:, :, h * 192:(h + 1) * 192].unsqueeze(1)  # (B,1,N,192) view
    k = qk[:, :, Cc + h * 192:Cc + (h + 1) * 192]  # (B,N,192) view
    out = torch.zeros(B, Npix, Npix, device="cuda", dtype=torch.float32)
    plan = ops.GemmPlan([(q, 1)], k, out, 1, out_ld=Npix, alpha=192 ** -0.5)
    for _ in range(3):
        plan.launch()
    torch.cuda.synchronize()
    ref = torch.bmm(q[:, 0].float(), k.float().transpose(1, 2)) * 192 ** -0.5
    assert rel_l2(out, ref) < 2e-5


@pytest.mark.parametrize("B,H,Cin,N,resid", [(2, 128, 64, 192, False), (3, 32, 192, 384, True), (5, 8, 192, 768, True),
                                              (2, 16, 128, 576, False), (1, 8, 192, 768, True), (1, 16, 64, 48, False)])
def test_fused_groupnorm_statistics(B, H, Cin, N, resid):
    """Epilogue-fused per-(sample, channel) [sum, sumsq] (int64, 2^20 fixed point) of the stored bf16 output."""
    ops = _setup()
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H)
    a = torch.randn(B, H, H, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, 9 * Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(B, H, H, N, device="cuda", generator=g).bfloat16() if resid else None
    out = torch.zeros(B, H, H, N, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.int64)
    plan = ops.GemmPlan([(a, 9)], w, out, 0, out_ld=N, bias=bias, resid=r, resid_ld=N, alpha=0.7071, stats=stats)
    plan.launch()
    torch.cuda.synchronize()
    ref = ref_conv([(a, 9)], w, bias, r, 0.7071)
    assert rel_l2(out.float().permute(0, 3, 1, 2), ref) < 4e-3
    o = out.double()
    assert rel_l2(stats[..., 0].double() / 2 ** 20, o.sum((1, 2))) < 1e-5
    assert rel_l2(stats[..., 1].double() / 2 ** 20, (o * o).sum((1, 2))) < 1e-5
    s1 = stats.clone()
    stats.zero_()
    plan.launch()
    torch.cuda.synchronize()
    assert torch.equal(s1, stats)  # integer atomics: bit-reproducible



@pytest.mark.parametrize("B,H,Cin,N,bn,cg", [
    (3, 32, 64, 192, 192, 1), (5, 16, 128, 384, 192, 2), (2, 128, 64, 192, 192, 2), (7, 16, 64, 96, 96, 1),
    (46, 32, 64, 64, 64, 2), (3, 64, 32, 256, 128, 1),
])
def test_gemm_fused_groupnorm_apply(B, H, Cin, N, bn, cg):
    """conv3x3 -> GroupNorm -> AdaGN affine -> SiLU in one launch (evc_gemm_desc.gn_ss): the epilogue holds each
    accumulator tile in tensor memory until the sample's statistics are complete.  Checked against torch fp32 and for
    bit-identical results over repeated launches (the per-sample tickets make the order irrelevant)."""
    ops = _setup()
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H)
    a = torch.randn(B, H, H, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, 9 * Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ss = torch.randn(2 * N, device="cuda", generator=g) * 0.3
    groups = min(N // 4, 32)
    outs = []
    for rep in range(3):
        out = torch.full((B, H, H, N), float("nan"), device="cuda", dtype=torch.bfloat16)
        stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.int64)
        ticket = torch.zeros(B, device="cuda", dtype=torch.int32)
        plan = ops.GemmPlan([(a, 9)], w, out, 0, out_ld=N, bias=bias, bn=bn, stats=stats, cta_group=cg,
                            gn=dict(ss=torch.zeros_like(ss), ticket=ticket, eps=1e-5, groups=groups, adagn=True))
        plan.launch(gn_ss=ss)
        torch.cuda.synchronize()
        outs.append(out)
    x = F.conv2d(a.float().permute(0, 3, 1, 2), w.float().reshape(N, 3, 3, Cin).permute(0, 3, 1, 2), padding=1)
    x = x + bias.view(1, -1, 1, 1)
    h = F.group_norm(x, groups, None, None, eps=1e-5)
    ref = F.silu(h * (1 + ss[:N]).view(1, -1, 1, 1) + ss[N:].view(1, -1, 1, 1))
    got = outs[0].float().permute(0, 3, 1, 2)
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < 5e-3, rel_l2(got, ref)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert int(ticket.min()) == int(ticket.max()) == (H * H // 128) * (N // bn)


def test_gemm_fused_groupnorm_guard_bands():
    """The fused GroupNorm-apply epilogue writes only `out`, `stats` and `gn_ticket` (guard bands around all three;
    an odd number of 128-row tiles, so the second CTA of the last pair has no tile)."""
    ops = _setup()
    B, H, Cin, N = 3, 16, 64, 192  # 6 tiles -> 3 pairs; B = 3 with 2 tiles per sample
    g = torch.Generator(device="cuda").manual_seed(77)
    a = torch.randn(B, H, H, Cin, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, 9 * Cin, device="cuda", generator=g) / (9 * Cin) ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ss = torch.randn(2 * N, device="cuda", generator=g) * 0.3
    pad = 1024
    obuf = torch.full((B * H * H * N + 2 * pad,), float("nan"), device="cuda", dtype=torch.bfloat16)
    out = obuf[pad:pad + B * H * H * N].view(B, H, H, N)
    sbuf = torch.zeros(B * N * 2 + 2 * pad, device="cuda", dtype=torch.int64)
    stats = sbuf[pad:pad + B * N * 2].view(B, N, 2)
    tbuf = torch.zeros(B + 2 * pad, device="cuda", dtype=torch.int32)
    ticket = tbuf[pad:pad + B]
    for cg in (1, 2):
        obuf.fill_(float("nan")); sbuf.zero_(); tbuf.zero_()
        plan = ops.GemmPlan([(a, 9)], w, out, 0, out_ld=N, bias=bias, bn=192, stats=stats, cta_group=cg,
                            gn=dict(ss=ss, ticket=ticket, eps=1e-5, groups=32, adagn=True))
        plan.launch()
        torch.cuda.synchronize()
        assert torch.isfinite(out.float()).all()
        assert bool(torch.isnan(obuf[:pad].float()).all()) and bool(torch.isnan(obuf[pad + B * H * H * N:].float()).all())
        assert int(sbuf[:pad].abs().sum()) == 0 and int(sbuf[pad + B * N * 2:].abs().sum()) == 0
        assert int(tbuf[:pad].abs().sum()) == 0 and int(tbuf[pad + B:].abs().sum()) == 0
        assert ticket.tolist() == [2, 2, 2]


def test_gemm_fused_groupnorm_rejects_oversized_samples():
    """ADVICE r01 / VERDICT item 8: a sample whose tiles do not fit in two rounds of the persistent grid would make a
    CTA wait for a tile it owns itself (e.g. image_size 256: 512 tiles per sample).  The plan must be refused on the
    host (EVC_ERR_UNSUPPORTED) -- the engine then keeps the unfused gn_apply launch -- and nothing may fault."""
    ops = _setup()
    from evcdiff._lib import EvcError
    from evcdiff.engine import EngineBase
    B, H, Cin, N = 1, 256, 64, 64
    a = torch.zeros(B, H, H, Cin, device="cuda", dtype=torch.bfloat16)
    w = torch.zeros(N, 9 * Cin, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(B, H, H, N, device="cuda", dtype=torch.bfloat16)
    stats = torch.zeros(B, N, 2, device="cuda", dtype=torch.int64)
    ticket = torch.zeros(B, device="cuda", dtype=torch.int32)
    gn = dict(ss=torch.zeros(2 * N, device="cuda"), ticket=ticket, eps=1e-5, groups=16, adagn=True)
    with pytest.raises(EvcError, match="unfused"):
        ops.GemmPlan([(a, 9)], w, out, 0, out_ld=N, bias=torch.zeros(N, device="cuda"), bn=64, stats=stats, gn=gn)
    # the engine-side mirror takes the same decision before any buffer is planned
    assert not ops.gn_fuse_fits(512, 512, 1)
    assert ops.gn_fuse_fits(128, 128 * 46, 1) and ops.gn_fuse_fits(128, 128, 1)
    eng = EngineBase.__new__(EngineBase)
    eng.split = False
    assert not eng.gn_fusable(1, 256, 256, [(64, 9)], 64)
    assert eng.gn_fusable(1, 128, 128, [(192, 9)], 192)
    # a fitting shape still works, and no tile wait ever gave up
    plan = ops.GemmPlan([(a[:, :128, :128].contiguous(), 9)], w, out[:, :128, :128].contiguous(), 0, out_ld=N,
                        bias=torch.zeros(N, device="cuda"), bn=64, stats=stats, gn=gn)
    plan.launch()
    torch.cuda.synchronize()
    assert ops.gemm_fault_count() == 0


SPLITK_CASES = [
    # name, B, H, W, [(C, taps)], N, out_mode, resid, alpha, bn, split_k, cta_group
    ("sk_w8_n768_s4", 6, 8, 8, [(768, 9)], 768, 0, True, 0.70710678, 96, 6, 1),
    ("sk_w8_skip_s3", 5, 8, 8, [(576, 9), (768, 1), (576, 1)], 768, 0, False, 0.70710678, 256, 3, 1),
    ("sk_w16_n576_s4", 3, 16, 16, [(384, 9)], 576, 0, False, 1.0, 192, 4, 1),
    ("sk_w16_pair_s2", 4, 16, 16, [(192, 9)], 384, 0, True, 1.0, 128, 2, 2),
    ("sk_w32_f32rows_s5", 1, 32, 32, [(192, 9)], 384, 1, False, 1.0, 64, 5, 1),
    ("sk_w8_f32T_s7", 1, 8, 8, [(576, 9)], 64, 3, False, 1.0, 32, 7, 1),
    ("sk_uneven_kb_s4", 2, 8, 8, [(192, 9), (64, 1)], 256, 0, False, 1.0, 128, 4, 1),  # 28 K blocks over 4 slices
    ("sk_w8_n768_s4_serial", 6, 8, 8, [(768, 9)], 768, 0, True, 0.70710678, 96, 6, 1),
    ("sk_w16_s3_serial", 3, 16, 16, [(384, 9)], 576, 0, False, 1.0, 192, 3, 1),
]


@pytest.mark.parametrize("case", SPLITK_CASES, ids=[c[0] for c in SPLITK_CASES])
def test_split_k_gemm(case):
    """Split-K (evc_gemm_desc.split_k): K slices of a tile on different CTAs, fp32 partial tiles in a workspace, the last
    arriver adds them in the fixed order 0..S-1 and runs the normal epilogue (bias, residual, alpha, TMA / per-thread
    stores, fused GroupNorm statistics).  Against torch fp32, against the unsplit launch (same values up to the fp32
    summation order: identical after bf16 rounding except for rare last-bit flips), and bit-identical run to run."""
    ops = _setup()
    name, B, H, W, segspec, N, out_mode, use_resid, alpha, bn, S, cg = case
    g = torch.Generator(device="cuda").manual_seed(hash(name) % 2**31)
    segs = [(torch.randn(B, H, W, Cc, device="cuda", generator=g).bfloat16(), taps) for Cc, taps in segspec]
    K = sum(Cc * taps for Cc, taps in segspec)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(B, H, W, N, device="cuda", generator=g).bfloat16() if use_resid else None
    ws = torch.empty(ops.SPLIT_K_WS_BYTES, dtype=torch.uint8, device="cuda")
    dt = torch.bfloat16 if out_mode == 0 else torch.float32
    shape = (B, H, W, N) if out_mode in (0, 1) else (B, N, H, W)
    ld, bs = (N, 0) if out_mode in (0, 1) else (H * W, N * H * W)
    want_stats = out_mode == 0 and (H * W) % 32 == 0
    res = {}
    for split in (1, S):
        outs = []
        for rep in range(3):
            out = torch.full(shape, float("nan"), device="cuda", dtype=dt)
            st = torch.zeros(B, N, 2, device="cuda", dtype=torch.int64) if want_stats else None
            # "_serial" cases cap the grid below the unit count: the K slices of a tile are then not all resident and the
            # last arriver adds them alone (the cooperative finish needs grid == units)
            plan = ops.GemmPlan(segs, w, out, out_mode, out_ld=ld, out_bs=bs, bias=bias, resid=resid, resid_ld=N, alpha=alpha,
                                bn=bn, stats=st, cta_group=cg, split_k=split, sk_ws=ws,
                                max_ctas=8 if name.endswith("_serial") else 0)
            assert plan.split_k == split
            plan.launch()
            plan.launch()  # tickets are left at zero by the last arriver: relaunch without any reset
            torch.cuda.synchronize()
            outs.append((out, st))
        assert torch.isfinite(outs[0][0].float()).all(), f"{name}: unwritten outputs (split_k={split})"
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0]), f"{name}: not reproducible"
        res[split] = outs[0]
    ref = ref_conv(segs, w, bias, resid, alpha)
    got = res[S][0].float()
    got = got.permute(0, 3, 1, 2) if out_mode in (0, 1) else got
    assert rel_l2(got, ref) < (4e-3 if out_mode == 0 else 2e-5), f"{name}: rel-L2 {rel_l2(got, ref):.3e}"
    base = res[1][0].float()
    base = base.permute(0, 3, 1, 2) if out_mode in (0, 1) else base
    # fp32 outputs differ by the summation order of the K slices only (a few ulp of the largest partial sum)
    assert rel_l2(got, base) < (2e-3 if out_mode == 0 else 1e-5), f"{name}: split vs unsplit {rel_l2(got, base):.3e}"
    if want_stats:
        # statistics of the stored bf16 values: each launch's statistics match its own output (two launches accumulated)
        o = res[S][0].float()
        s_ref = torch.stack([o.sum(dim=(1, 2)), (o * o).sum(dim=(1, 2))], -1) * 2
        s_got = res[S][1].double() / 2 ** 20
        assert rel_l2(s_got, s_ref) < 1e-4, f"{name}: fused statistics {rel_l2(s_got, s_ref):.3e}"
    assert ops.gemm_fault_count() == 0
