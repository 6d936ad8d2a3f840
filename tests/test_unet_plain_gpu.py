"""GPU parity of the models/unet.py variant (BASELINE config 5) against the reference goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import samplers as S
from oracle import unet_plain as U

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def T(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def build(cfg, seed):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from evcdiff.models.unet import UNet_DDPM
    net = UNet_DDPM(cfg)
    sd = common.seeded_state_dict(U.unet_param_shapes(cfg), seed=seed, active=True)
    r = net.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys and set(r.missing_keys) <= {"betas", "alphas", "alphas_prev"}
    return net.to(DEV).eval(), {k: v.to(DEV) for k, v in sd.items()}


def _tape(seed, n, shape):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(shape, generator=g) for _ in range(n)]


def _report(net, sd, cfg, x, lab, cond):
    taps = {}
    U.unet_forward(sd, cfg, x, lab, cond, taps=taps)
    eng = net.engine(x.shape[0], DEV)
    return [(k, round(common.rel_l2(a.t.float().permute(0, 3, 1, 2), taps[k]), 4)) for k, a in eng.taps.items() if k in taps]


@pytest.mark.parametrize("mode", ["deep", "deeper"])
def test_eps_and_ddpm_vs_golden(mode):
    from evcdiff import models as M
    g = dict(np.load(os.path.join(G, "unet_plain.npz")))
    cfg = common.make_config(ngf=32, image_size=16, device=DEV)
    cfg.mode = mode
    net, sd = build(cfg, 21)
    x, cond = T(g[f"{mode}_x"]), T(g[f"{mode}_cond"])
    for lab in (0, 990):
        labels = torch.full((2,), lab, dtype=torch.long, device=DEV)
        eps = net(x, labels, cond=cond)
        err = common.rel_l2(eps, T(g[f"{mode}_eps_{lab}"]))
        assert err < 3e-2, (err, _report(net, sd, cfg, x, labels, cond))
    tape = _tape(23, 9, x.shape)
    y = M.ddpm_sampler(x.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=10, noise=tape)
    assert common.rel_l2(y, T(g[f"{mode}_ddpm10"])) < 2.5e-2


def test_mid_size_vs_oracle():
    """ngf=64, 64x64 ('deep'): stride-2 convs, nearest upsampling, attention at 32x32 (N=1024) and 8x8."""
    from evcdiff import models as M
    cfg = common.make_config(ngf=64, image_size=64, device=DEV)
    cfg.mode = "deep"
    net, sd = build(cfg, 31)
    g = torch.Generator(device=DEV).manual_seed(32)
    x = torch.randn(3, 15, 64, 64, device=DEV, generator=g)
    cond = torch.rand(3, 6, 64, 64, device=DEV, generator=g) * 2 - 1
    labels = torch.full((3,), 300, dtype=torch.long, device=DEV)
    eps = net(x, labels, cond=cond)
    ref = U.unet_forward(sd, cfg, x, labels, cond)
    err = common.rel_l2(eps, ref)
    assert err < 3e-2, (err, _report(net, sd, cfg, x, labels, cond))
    # sampler arithmetic through the graph-captured loop (FPNDM) against the oracle sampler on the same network
    model = lambda xx, yy: net(xx, yy, cond=cond)
    y = M.FPNDM_sampler(x.clone(), net, cond=cond, final_only=True, subsample_steps=10)
    r = S.fpndm_sampler(x.clone(), model, (net.betas, net.alphas, net.alphas_prev), subsample_steps=10)
    assert common.rel_l2(y[0], r) < 2e-3


def test_config5_shape_deep_full():
    """The BASELINE config-5 network itself: ngf=192, 128x128, 'deep' (80.4 M parameters; attention over 4096 tokens)."""
    cfg = common.make_config(device=DEV)
    cfg.mode = "deep"
    net, sd = build(cfg, 41)
    g = torch.Generator(device=DEV).manual_seed(42)
    x = torch.randn(1, 15, 128, 128, device=DEV, generator=g)
    cond = torch.rand(1, 6, 128, 128, device=DEV, generator=g) * 2 - 1
    labels = torch.full((1,), 500, dtype=torch.long, device=DEV)
    eps = net(x, labels, cond=cond)
    ref = U.unet_forward(sd, cfg, x, labels, cond)
    err = common.rel_l2(eps, ref)
    assert err < 3e-2, (err, _report(net, sd, cfg, x, labels, cond))
