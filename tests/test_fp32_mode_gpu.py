"""GPU parity of the split-precision ("fp32-tolerance") mode: north_star asks for per-step x_t within rel-L2 1e-3 of
the fp32 reference (1e-2 in bf16).  Every bf16 tensor is a (hi, lo) pair (16 mantissa bits) and every product is
hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import common
from oracle import ncsnpp as O
from oracle import samplers as S

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def T(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def split(x):
    hi = x.to(torch.bfloat16)
    return hi.contiguous(), (x - hi.float()).to(torch.bfloat16).contiguous()


@pytest.mark.parametrize("B,H,Cin,N,resid,taps", [(2, 32, 192, 384, True, 9), (1, 128, 64, 192, False, 9),
                                                   (3, 16, 384, 576, True, 1), (5, 8, 192, 768, False, 9)])
def test_split_gemm(B, H, Cin, N, resid, taps):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from evcdiff import ops
    g = torch.Generator(device=DEV).manual_seed(B * 100 + H)
    a = torch.randn(B, H, H, Cin, device=DEV, generator=g)
    w = torch.randn(N, taps * Cin, device=DEV, generator=g) / (taps * Cin) ** 0.5
    bias = torch.randn(N, device=DEV, generator=g)
    r = torch.randn(B, H, H, N, device=DEV, generator=g) if resid else None
    a_hi, a_lo = split(a)
    w_hi, w_lo = split(w)
    r_hi, r_lo = split(r) if resid else (None, None)
    out = torch.zeros(B, H, H, N, device=DEV, dtype=torch.bfloat16)
    out_lo = torch.zeros_like(out)
    stats = torch.zeros(B, N, 2, device=DEV, dtype=torch.int64)
    plan = ops.GemmPlan([(a_hi, taps)], w_hi, out, 0, out_ld=N, bias=bias, resid=r_hi, resid_ld=N, alpha=0.7071,
                        stats=stats, segs_lo=[a_lo], w_lo=w_lo, out_lo=out_lo, resid_lo=r_lo)
    plan.launch()
    torch.cuda.synchronize()
    x = a.permute(0, 3, 1, 2).double()
    if taps == 9:
        ref = F.conv2d(x, w.double().reshape(N, 3, 3, Cin).permute(0, 3, 1, 2), padding=1)
    else:
        ref = F.conv2d(x, w.double().reshape(N, Cin, 1, 1))
    ref = ref + bias.double().view(1, -1, 1, 1)
    if resid:
        ref = ref + r.double().permute(0, 3, 1, 2)
    ref = ref * 0.7071
    got = (out.float() + out_lo.float()).permute(0, 3, 1, 2)
    err = common.rel_l2(got, ref)
    assert err < 3e-5, err  # bf16 mode: ~3e-3
    o = (out.double() + out_lo.double())
    assert common.rel_l2(stats[..., 0].double() / 2 ** 20, o.sum((1, 2))) < 1e-5


def _build(cfgf, seed, active=True):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    cfg = cfgf(device=DEV)
    net = UNetMore_DDPM(cfg)
    sd = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=seed, active=active)
    net.load_state_dict(sd, strict=False)
    net = net.to(DEV).eval()
    net.precision = "fp32"
    return cfg, net, {k: v.to(DEV) for k, v in sd.items()}


@pytest.mark.parametrize("tag,cfgf,seed", [("tiny_act", common.tiny_config, 1), ("gpu64", common.gpu64_config, 4)])
def test_eps_fp32_mode_vs_golden(tag, cfgf, seed):
    small = dict(np.load(os.path.join(G, "ncsnpp_small.npz")))
    cfg, net, sd = _build(cfgf, seed)
    x, cond = T(small[f"{tag}_x"]), T(small[f"{tag}_cond"])
    for lab in (0, 990):
        eps = net(x, torch.full((2,), lab, dtype=torch.long, device=DEV), cond=cond)
        err = common.rel_l2(eps, T(small[f"{tag}_eps_{lab}"]))
        assert err < 1e-3, (tag, lab, err)
    net.precision = "bf16"  # both engines coexist (keyed by precision)
    eps16 = net(x, torch.full((2,), 0, dtype=torch.long, device=DEV), cond=cond)
    assert 1e-3 < common.rel_l2(eps16, T(small[f"{tag}_eps_0"])) < 3e-2


def test_eps_fp32_mode_full_model():
    full = dict(np.load(os.path.join(G, "ncsnpp_full.npz")))
    cfg, net, sd = _build(common.full_config, 9)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(1, 15, 128, 128, generator=g).to(DEV)
    cond = (torch.rand(1, 6, 128, 128, generator=g, dtype=torch.float64) * 2 - 1).to(DEV)
    eps = net(x, torch.full((1,), 990, dtype=torch.long, device=DEV), cond=cond)
    err = common.rel_l2(eps[:, :, ::4, ::4], T(full["full_eps_990_sub4"]))
    assert err < 1e-3, err


def test_ddpm_trajectory_fp32_mode():
    """Per-step x_t within 1e-3 of the fp32 oracle over a 50-step DDPM run (same noise tape), graph-captured FPNDM runs."""
    from evcdiff import models as M
    cfg, net, sd = _build(common.gpu64_config, 4)
    g = torch.Generator(device=DEV).manual_seed(15)
    x_T = torch.randn(2, 15, 32, 32, device=DEV, generator=g)
    cond = torch.rand(2, 6, 32, 32, device=DEV, generator=g, dtype=torch.float64) * 2 - 1
    sched = S.schedule(cfg, DEV)
    model = lambda x, y: O.ncsnpp_forward(sd, cfg, x, y, cond)
    gt = torch.Generator().manual_seed(78)
    tape = [torch.randn(x_T.shape, generator=gt).to(DEV) for _ in range(49)]
    imgs = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=False, denoise=True, subsample_steps=50, noise=tape,
                          precision="fp32")
    trace = []
    S.ddpm_sampler(x_T.clone(), model, sched, lambda i: tape[i], subsample_steps=50, trace=trace)
    errs = [common.rel_l2(imgs[i].to(DEV), trace[i][0]) for i in range(51)]
    assert max(errs) < 1e-3, (max(errs), errs[::10])
    y = M.FPNDM_sampler(x_T.clone(), net, cond=cond, final_only=True, subsample_steps=10, precision="fp32")
    assert torch.isfinite(y).all()
