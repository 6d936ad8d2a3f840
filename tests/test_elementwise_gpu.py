"""GPU parity of the HBM-bound kernels (GroupNorm stats/apply, FIR, softmax, pack, sampler updates, temb)
against the oracle / plain torch fp32.  bf16 outputs: rel-L2 <= 4e-3 (one rounding); fp32 outputs <= 1e-6."""
import math

import pytest
import torch
import torch.nn.functional as F

import common
from oracle import ncsnpp as O
from oracle import samplers as S

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    from evcdiff import ops
    return ops


def nhwc(x):  # NCHW fp32 -> NHWC bf16
    return x.permute(0, 2, 3, 1).contiguous().bfloat16()


def nchw(x):  # NHWC bf16 -> NCHW fp32
    return x.float().permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("B,H,C0,C1,adagn,silu,eps", [
    (2, 16, 192, 0, True, True, 1e-5), (3, 8, 768, 576, True, True, 1e-5), (2, 32, 384, 192, True, True, 1e-5),
    (2, 16, 576, 0, False, False, 1e-6), (1, 128, 192, 0, False, True, 1e-5), (2, 4, 64, 32, True, True, 1e-5),
])
def test_groupnorm(B, H, C0, C1, adagn, silu, eps):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(1)
    C = C0 + C1
    x0 = nhwc(torch.randn(B, C0, H, H, device=DEV, generator=g) * 1.7 + 0.3)
    x1 = nhwc(torch.randn(B, C1, H, H, device=DEV, generator=g) * 0.6 - 0.2) if C1 else None
    ss = torch.randn(2 * C, device=DEV, generator=g) * 0.3
    if not adagn:
        ss[:C] += 1.0
    st0 = torch.zeros(B, C0, 2, device=DEV, dtype=torch.int64)
    ops.gn_stats(x0, B, H * H, C0, st0)
    st1 = None
    if C1:
        st1 = torch.zeros(B, C1, 2, device=DEV, dtype=torch.int64)
        ops.gn_stats(x1, B, H * H, C1, st1)
    y = torch.empty(B, H, H, C, device=DEV, dtype=torch.bfloat16)
    groups = O.gn_groups(C)
    ops.gn_apply(x0, C0, x1, C1, B, H * H, st0, st1, groups, eps, ss, adagn, silu, y)
    torch.cuda.synchronize()
    xc = torch.cat([nchw(x0)] + ([nchw(x1)] if C1 else []), 1)
    ref_sum = xc.sum((2, 3))
    assert common.rel_l2(st0[..., 0].double() / 2 ** 20, ref_sum[:, :C0]) < 1e-5
    assert common.rel_l2(st0[..., 1].double() / 2 ** 20, (xc * xc).sum((2, 3))[:, :C0]) < 1e-5
    h = F.group_norm(xc, groups, None, None, eps=eps)
    gam = (1 + ss[:C]) if adagn else ss[:C]
    h = h * gam.view(1, C, 1, 1) + ss[C:].view(1, C, 1, 1)
    if silu:
        h = F.silu(h)
    assert common.rel_l2(nchw(y), h) < 4e-3


@pytest.mark.parametrize("B,H,C", [(2, 8, 192), (1, 64, 192), (3, 16, 576), (2, 4, 64)])
def test_fir(B, H, C):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(2)
    x = nhwc(torch.randn(B, C, H, H, device=DEV, generator=g))
    up = torch.empty(B, 2 * H, 2 * H, C, device=DEV, dtype=torch.bfloat16)
    dn = torch.empty(B, H // 2, H // 2, C, device=DEV, dtype=torch.bfloat16)
    ops.fir_resample(x, up, B, H, H, C, True)
    ops.fir_resample(x, dn, B, H, H, C, False)
    torch.cuda.synchronize()
    assert common.rel_l2(nchw(up), O.fir_up2(nchw(x))) < 4e-3
    assert common.rel_l2(nchw(dn), O.fir_down2(nchw(x))) < 4e-3


@pytest.mark.parametrize("B,H,C0,C1,up", [
    (2, 16, 192, 0, True), (2, 16, 192, 0, False), (3, 8, 768, 576, True), (2, 32, 384, 192, True),
    (2, 32, 384, 192, False), (1, 64, 192, 0, False), (1, 64, 192, 192, True), (2, 4, 64, 32, True), (2, 4, 64, 32, False),
    (1, 34, 192, 0, False), (1, 20, 64, 0, True), (2, 2, 96, 0, False), (2, 1, 128, 96, True), (2, 2, 96, 64, True),
    (1, 128, 192, 0, False), (2, 16, 576, 0, False), (1, 12, 24, 8, False), (3, 6, 40, 0, False), (1, 70, 64, 0, False),
])
def test_gn_fir_fused(B, H, C0, C1, up):
    """FIR(SiLU(AdaGN([x0|x1]))) and FIR(x) from one kernel (up / down res-block prologue, layerspp.py:598-611)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(4)
    C = C0 + C1
    x0 = nhwc(torch.randn(B, C0, H, H, device=DEV, generator=g) * 1.7 + 0.3)
    x1 = nhwc(torch.randn(B, C1, H, H, device=DEV, generator=g) * 0.6 - 0.2) if C1 else None
    ss = torch.randn(2 * C, device=DEV, generator=g) * 0.3
    st0 = torch.zeros(B, C0, 2, device=DEV, dtype=torch.int64)
    ops.gn_stats(x0, B, H * H, C0, st0)
    st1 = None
    if C1:
        st1 = torch.zeros(B, C1, 2, device=DEV, dtype=torch.int64)
        ops.gn_stats(x1, B, H * H, C1, st1)
    H2 = 2 * H if up else H // 2
    ya = torch.full((B, H2, H2, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    r0 = torch.full((B, H2, H2, C0), float("nan"), device=DEV, dtype=torch.bfloat16)
    r1 = torch.full((B, H2, H2, C1), float("nan"), device=DEV, dtype=torch.bfloat16) if C1 else None
    groups = O.gn_groups(C)
    ops.gn_fir(x0, C0, x1, C1, B, H, H, st0, st1, groups, 1e-5, ss, True, up, ya, r0, r1)
    torch.cuda.synchronize()
    xc = torch.cat([nchw(x0)] + ([nchw(x1)] if C1 else []), 1)
    h = F.group_norm(xc, groups, None, None, eps=1e-5)
    h = F.silu(h * (1 + ss[:C]).view(1, C, 1, 1) + ss[C:].view(1, C, 1, 1))
    fir = O.fir_up2 if up else O.fir_down2
    assert common.rel_l2(nchw(ya), fir(h)) < 4e-3
    assert common.rel_l2(nchw(r0), fir(nchw(x0))) < 4e-3
    if C1:
        assert common.rel_l2(nchw(r1), fir(nchw(x1))) < 4e-3


def test_nearest_up():
    ops = _ops()
    x = nhwc(torch.randn(2, 64, 8, 8, device=DEV))
    y = torch.empty(2, 16, 16, 64, device=DEV, dtype=torch.bfloat16)
    ops.nearest_up2(x, y, 2, 8, 8, 64)
    torch.cuda.synchronize()
    assert torch.equal(nchw(y), F.interpolate(nchw(x), scale_factor=2, mode="nearest"))


@pytest.mark.parametrize("rows,cols", [(64, 64), (1000, 256), (2048, 1024), (512, 4096), (32, 16), (8, 4)])
def test_softmax(rows, cols):
    ops = _ops()
    s = torch.randn(rows, cols, device=DEV) * 4
    p = torch.empty(rows, cols, device=DEV, dtype=torch.bfloat16)
    ops.softmax_rows(s, p, rows, cols)
    torch.cuda.synchronize()
    assert common.rel_l2(p.float(), F.softmax(s, -1)) < 4e-3


def test_pack_and_inverse():
    ops = _ops()
    x = torch.randn(3, 15, 16, 16, device=DEV)
    c = torch.rand(3, 6, 16, 16, device=DEV, dtype=torch.float64)
    dst = torch.zeros(3, 16, 16, 64, device=DEV, dtype=torch.bfloat16)
    ops.pack_nchw(x, dst, 0)
    ops.pack_nchw(c, dst, 15, scale=2.0, shift=-1.0)
    torch.cuda.synchronize()
    ref = torch.cat([x, (2 * c - 1).float()], 1).bfloat16().float()
    assert torch.equal(nchw(dst)[:, :21], ref)
    assert float(nchw(dst)[:, 21:].abs().max()) == 0.0
    fr = torch.empty_like(x)
    ops.inverse_transform(x, fr)
    assert common.rel_l2(fr, torch.clamp((x + 1) / 2, 0, 1)) < 1e-7


def test_temb_and_linear():
    ops = _ops()
    t = torch.tensor([0.0, 10.0, 990.0, 999.0, -0.5, 25.0, -1.0], device=DEV)
    from evcdiff.models.better.layers import get_timestep_embedding
    e = get_timestep_embedding(t, 192)
    ref = O.timestep_embedding(t, 192)
    assert float((e - ref).abs().max()) < 2e-5  # sinf/cosf of arguments up to ~1e3 in fp32
    g = torch.Generator(device=DEV).manual_seed(3)
    x = torch.randn(29, 768, device=DEV, generator=g)
    W = torch.randn(1000, 768, device=DEV, generator=g) / 28
    b = torch.randn(1000, device=DEV, generator=g)
    y = torch.empty(29, 1000, device=DEV)
    ops.linear_f32(x, W, b, y, act_in=True)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = F.linear(F.silu(x).double(), W.double(), b.double()).float()
    assert common.rel_l2(y, ref) < 1e-6


def test_sampler_updates_match_oracle_arithmetic():
    """DDPM / DDIM / denoise / PNDM transfer updates against the oracle's tensor expressions (fp32, CUDA)."""
    ops = _ops()
    from evcdiff._lib import PndmCoef, StepCoef
    cfg = common.tiny_config(device=DEV)
    betas, alphas, alphas_prev = S.schedule(cfg, DEV)
    steps, a, ap, b = S._subsample(alphas, alphas_prev, betas, 100)
    g = torch.Generator(device=DEV).manual_seed(4)
    x = torch.randn(2, 15, 16, 16, device=DEV, generator=g)
    e = torch.randn(2, 15, 16, 16, device=DEV, generator=g)
    nz = torch.randn(2, 15, 16, 16, device=DEV, generator=g)
    xin = torch.full((2, 16, 16, 64), 7.0, device=DEV, dtype=torch.bfloat16)
    for i in (0, 37, 98):
        ca, cap, cb = a[i], ap[i], b[i]
        x0 = ((1 / ca.sqrt()) * (x - (1 - ca).sqrt() * e)).clip_(-1, 1)
        ref = (cap.sqrt() * cb / (1 - ca)) * x0 + ((1 - cb).sqrt() * (1 - cap) / (1 - ca)) * x
        ref = ref + ((1 - cap) / (1 - ca) * cb).sqrt() * nz
        c = StepCoef(0, 1, float(1 / ca.sqrt()), float((1 - ca).sqrt()), float(cap.sqrt() * cb / (1 - ca)),
                     float((1 - cb).sqrt() * (1 - cap) / (1 - ca)), 0.0, float(((1 - cap) / (1 - ca) * cb).sqrt()))
        out = torch.empty_like(x)
        ops.sampler_update(x, e, nz, out, xin, c)
        torch.cuda.synchronize()
        assert common.rel_l2(out, ref) < 1e-6
        assert torch.equal(nchw(xin)[:, :15], out.bfloat16().float())
        # whole 16-byte chunks: the pad channel 15 is written as zero, the conditioning channels (16..) are untouched
        assert bool((nchw(xin)[:, 15] == 0).all()) and bool((nchw(xin)[:, 16:] == 7.0).all())
        # DDIM
        ref = cap.sqrt() * x0 + (1 - cap).sqrt() * e
        c = StepCoef(0, 1, float(1 / ca.sqrt()), float((1 - ca).sqrt()), float(cap.sqrt()), 0.0, float((1 - cap).sqrt()), 0.0)
        ops.sampler_update(x, e, None, out, None, c)
        torch.cuda.synchronize()
        assert common.rel_l2(out, ref) < 1e-6
    ref = x - (1 - a[-1]).sqrt() * e
    ops.sampler_update(x, e, None, out, None, StepCoef(1, 0, 0.0, float((1 - a[-1]).sqrt()), 0, 0, 0, 0))
    torch.cuda.synchronize()
    assert common.rel_l2(out, ref) < 1e-7
    # PNDM transfer + linear multistep through the public pndm mirror
    from evcdiff.models import pndm
    a_old = alphas.flip(0)
    for (t, tn) in [(50.0, 25.0), (0.0, -0.5), (950.0, 900.0)]:
        tt = torch.full((2,), t, device=DEV)
        tnn = torch.full((2,), tn, device=DEV)
        for clip in (True, False):
            ref = S.transfer(x, tt, tnn, e, a_old, clip_before=clip)
            got = pndm.transfer(x, tt, tnn, e, a_old, clip_before=clip)
            torch.cuda.synchronize()
            assert common.rel_l2(got, ref) < 1e-6
    es = [torch.randn_like(x) for _ in range(4)]
    ref = (1 / 24) * (55 * es[0] - 59 * es[1] + 37 * es[2] - 9 * es[3])
    got = pndm._combine(es, (55.0, -59.0, 37.0, -9.0), 1 / 24)
    torch.cuda.synchronize()
    assert common.rel_l2(got, ref) < 1e-6


def test_sampler_update_vector_and_scalar_paths_agree():
    """The float4 / 4-pixels-per-thread kernels (HW % 4 == 0, aligned) and the scalar fallback (ragged HW, misaligned
    views) compute bit-identical updates and write the same UNet-input rows; C = 15 and C = 9 (CP = 16)."""
    ops = _ops()
    from evcdiff._lib import PndmCoef, StepCoef
    g = torch.Generator(device=DEV).manual_seed(41)
    c = StepCoef(0, 1, 1.7, 0.6, 0.3, 0.65, 0.0, 0.2)
    pc = PndmCoef()
    pc.n_e, pc.clip, pc.w_scale, pc.d, pc.p, pc.q = 4, 1, 1 / 24, 0.05, 0.9, 1.3
    for j, w in enumerate((55.0, -59.0, 37.0, -9.0)):
        pc.w[j] = w
    for C in (15, 9):
        B, H, W = 3, 12, 12
        al = [torch.randn(B, C, H, W, device=DEV, generator=g) for _ in range(7)]  # aligned planes -> vector path
        x, e, nz = al[0], al[1], al[2]
        xin_v = torch.full((B, H, W, 64), 3.0, device=DEV, dtype=torch.bfloat16)
        out_v = torch.empty_like(x)
        ops.sampler_update(x, e, nz, out_v, xin_v, c)
        # misaligned copies of the same data (offset by one float) force the scalar kernel
        xm, em, nm = [torch.empty(B * C * H * W + 1, device=DEV)[1:].view(B, C, H, W) for _ in range(3)]
        xm.copy_(x); em.copy_(e); nm.copy_(nz)
        xin_s = torch.full((B, H, W, 64), 3.0, device=DEV, dtype=torch.bfloat16)
        out_s = torch.empty_like(x)
        ops.sampler_update(xm, em, nm, out_s, xin_s, c)
        torch.cuda.synchronize()
        assert torch.equal(out_v, out_s) and torch.equal(xin_v, xin_s)
        CP = (C + 7) // 8 * 8
        assert bool((xin_v[..., C:CP] == 0).all()) and bool((xin_v[..., CP:] == 3.0).all())
        assert torch.equal(xin_v[..., :C].float(), out_v.permute(0, 2, 3, 1).bfloat16().float())
        # PNDM: 4-term multistep, et written out, in-place x
        es = [al[3], al[4], al[5], al[6]]
        xv, xs = x.clone(), xm.clone()
        etv, ets = torch.empty_like(x), torch.empty_like(x)
        ops.pndm_update(xv, es, xv, etv, xin_v, pc)
        esm_v = [torch.empty(B * C * H * W + 1, device=DEV)[1:].view(B, C, H, W) for _ in range(4)]
        for a_, b_ in zip(esm_v, es):
            a_.copy_(b_)
        ops.pndm_update(xs, esm_v, xs, ets, xin_s, pc)
        torch.cuda.synchronize()
        assert torch.equal(xv, xs) and torch.equal(etv, ets) and torch.equal(xin_v, xin_s)
        ref_et = (1 / 24) * (55 * es[0] - 59 * es[1] + 37 * es[2] - 9 * es[3])
        assert common.rel_l2(etv, ref_et) < 1e-6
    # ragged H*W (not a multiple of 4): scalar path, against the tensor expression
    x = torch.randn(2, 15, 3, 3, device=DEV, generator=g)
    e = torch.randn(2, 15, 3, 3, device=DEV, generator=g)
    out = torch.empty_like(x)
    ops.sampler_update(x, e, None, out, None, StepCoef(1, 0, 0.0, 0.25, 0, 0, 0, 0))
    torch.cuda.synchronize()
    assert common.rel_l2(out, x - 0.25 * e) < 1e-7


def test_frames_to_uint8_and_final_conv_redirect():
    """uint8 frame format for the end-of-run gather (round half to even like torch.round) and the eps_out redirect of
    the last conv (F-PNDM writes eps straight into its history ring)."""
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(43)
    fr = torch.rand(3, 5, 3, 16, 16, device=DEV, generator=g)
    fr.view(-1)[:8] = torch.tensor([0.0, 1.0, 0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 0.999, 0.001], device=DEV)
    out = torch.empty(fr.shape, dtype=torch.uint8, device=DEV)
    ops.frames_to_uint8(fr, out)
    torch.cuda.synchronize()
    assert torch.equal(out, (fr * 255.0).round().clamp(0, 255).to(torch.uint8))
    odd = torch.rand(1027, device=DEV, generator=g)
    o2 = torch.empty(1027, dtype=torch.uint8, device=DEV)
    ops.frames_to_uint8(odd, o2)
    assert torch.equal(o2, (odd * 255.0).round().to(torch.uint8))
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    cfg = common.gpu64_config(device=DEV)
    net = UNetMore_DDPM(cfg).to(DEV).eval()
    eng = net.engine(2, DEV)
    eng.set_labels([100.0])
    x = torch.randn(2, 15, 32, 32, device=DEV, generator=g)
    eng.load_input(x, None)
    a = eng.forward(0).clone()
    dst = torch.full_like(a, float("nan"))
    b = eng.forward(0, eps_out=dst)
    torch.cuda.synchronize()
    assert b is dst and torch.equal(a, dst)


def _guarded(shape, dtype, fill):
    """A tensor placed in the middle of a larger allocation; returns (view, check) where check() asserts the guard
    bands are untouched (compute-sanitizer is not available on the GPU pool, so out-of-bounds stores are caught here)."""
    n = int(torch.tensor(shape).prod())
    pad = 4096
    buf = torch.full((n + 2 * pad,), fill, device=DEV, dtype=dtype)
    view = buf[pad:pad + n].view(*shape)

    def check():
        lo, hi = buf[:pad], buf[pad + n:]
        ok = (lo != lo).all() and (hi != hi).all() if fill != fill else bool((lo == fill).all() and (hi == fill).all())
        assert bool(ok), "store outside the output tensor"
    return view, check


@pytest.mark.parametrize("B,H,C0,C1,up", [(2, 16, 192, 0, True), (3, 16, 192, 96, False), (2, 6, 64, 0, True), (1, 10, 96, 32, False)])
def test_gn_fir_stays_inside_its_outputs(B, H, C0, C1, up):
    ops = _ops()
    g = torch.Generator(device=DEV).manual_seed(9)
    C = C0 + C1
    x0 = nhwc(torch.randn(B, C0, H, H, device=DEV, generator=g))
    x1 = nhwc(torch.randn(B, C1, H, H, device=DEV, generator=g)) if C1 else None
    ss = torch.randn(2 * C, device=DEV, generator=g) * 0.3
    st0 = torch.zeros(B, C0, 2, device=DEV, dtype=torch.int64)
    ops.gn_stats(x0, B, H * H, C0, st0)
    st1 = None
    if C1:
        st1 = torch.zeros(B, C1, 2, device=DEV, dtype=torch.int64)
        ops.gn_stats(x1, B, H * H, C1, st1)
    H2 = 2 * H if up else H // 2
    nan = float("nan")
    ya, ca = _guarded((B, H2, H2, C), torch.bfloat16, nan)
    r0, c0 = _guarded((B, H2, H2, C0), torch.bfloat16, nan)
    r1, c1 = _guarded((B, H2, H2, C1), torch.bfloat16, nan) if C1 else (None, None)
    ops.gn_fir(x0, C0, x1, C1, B, H, H, st0, st1, O.gn_groups(C), 1e-5, ss, True, up, ya, r0, r1)
    torch.cuda.synchronize()
    for t, chk in ((ya, ca), (r0, c0), (r1, c1)):
        if t is not None:
            assert torch.isfinite(t.float()).all()
            chk()


def test_gn_apply_stays_inside_its_output():
    ops = _ops()
    B, H, C0, C1 = 3, 10, 96, 32
    g = torch.Generator(device=DEV).manual_seed(10)
    x0 = nhwc(torch.randn(B, C0, H, H, device=DEV, generator=g))
    x1 = nhwc(torch.randn(B, C1, H, H, device=DEV, generator=g))
    ss = torch.randn(2 * (C0 + C1), device=DEV, generator=g) * 0.3
    st0 = torch.zeros(B, C0, 2, device=DEV, dtype=torch.int64)
    st1 = torch.zeros(B, C1, 2, device=DEV, dtype=torch.int64)
    ops.gn_stats(x0, B, H * H, C0, st0)
    ops.gn_stats(x1, B, H * H, C1, st1)
    y, chk = _guarded((B, H, H, C0 + C1), torch.bfloat16, float("nan"))
    ops.gn_apply(x0, C0, x1, C1, B, H * H, st0, st1, O.gn_groups(C0 + C1), 1e-5, ss, True, True, y)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()
    chk()
