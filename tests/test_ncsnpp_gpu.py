"""GPU parity of the whole NCSN++ evaluation and of the samplers against the oracle and the committed golden
vectors (generated from the unmodified reference).

Tolerances (north_star, bf16 compute mode): per-step x_t rel-L2 <= 1e-2; final-frame PSNR within 0.05 dB of the
fp32 result.  eps of a single evaluation is compared directly as well (SURVEY.md section 4 trap: at default
init x_t is insensitive to the UNet), bound 1.5e-2 (measured 0.7-1.3e-2; the reference itself under bf16 autocast
measures 1.9e-2, section 4a).
"""
import os

import numpy as np
import pytest
import torch

import common
from oracle import ncsnpp as O
from oracle import samplers as S

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EPS_TOL = 1.5e-2
XT_TOL = 1e-2


def T(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


def build(cfgf, seed, active=True):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    cfg = cfgf(device=DEV)
    net = UNetMore_DDPM(cfg)
    sd = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=seed, active=active)
    missing = net.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    assert set(missing.missing_keys) <= {"betas", "alphas", "alphas_prev", "unet.sigmas"}
    net = net.to(DEV).eval()
    sd = {k: v.to(DEV) for k, v in sd.items()}
    return cfg, net, sd


@pytest.fixture(scope="module")
def small():
    return dict(np.load(os.path.join(G, "ncsnpp_small.npz")))


def _layer_report(net, sd, cfg, x, lab, cond):
    taps = {}
    ref = O.ncsnpp_forward(sd, cfg, x, lab, cond, taps=taps)
    eng = net.engine(x.shape[0], DEV)
    rep = []
    for name, act in eng.taps.items():
        if name in taps and act.t.shape[0] == x.shape[0]:
            got = act.t.float().permute(0, 3, 1, 2)
            rep.append((name, common.rel_l2(got, taps[name])))
    return ref, rep


@pytest.mark.parametrize("tag,cfgf,seed", [("tiny_act", common.tiny_config, 1), ("gpu64", common.gpu64_config, 4)])
def test_eps_vs_golden(small, tag, cfgf, seed):
    cfg, net, sd = build(cfgf, seed)
    x, cond = T(small[f"{tag}_x"]), T(small[f"{tag}_cond"])
    for lab in (0, 990):
        labels = torch.full((2,), lab, dtype=torch.long, device=DEV)
        eps = net(x, labels, cond=cond)
        torch.cuda.synchronize()
        err = common.rel_l2(eps, T(small[f"{tag}_eps_{lab}"]))
        if err >= EPS_TOL:
            _, rep = _layer_report(net, sd, cfg, x, labels, cond)
            pytest.fail(f"{tag} label {lab}: eps rel-L2 {err:.3e}; per-module (persistent taps only): {rep}")


def test_eps_fractional_label(small):
    cfg, net, sd = build(common.tiny_config, 1)
    x, cond = T(small["tiny_act_x"]), T(small["tiny_act_cond"])
    eps = net(x, torch.full((2,), -0.5, device=DEV), cond=cond)
    assert common.rel_l2(eps, T(small["tiny_act_eps_m0p5"])) < EPS_TOL


def test_eps_full_model():
    """configs/mine.yml model (262 M parameters, 128x128), B=1, against the reference golden (sub-sampled)."""
    full = dict(np.load(os.path.join(G, "ncsnpp_full.npz")))
    cfg, net, sd = build(common.full_config, 9)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(1, 15, 128, 128, generator=g).to(DEV)
    cond = (torch.rand(1, 6, 128, 128, generator=g, dtype=torch.float64) * 2 - 1).to(DEV)
    for lab in (0, 990):
        labels = torch.full((1,), lab, dtype=torch.long, device=DEV)
        eps = net(x, labels, cond=cond)
        torch.cuda.synchronize()
        err = common.rel_l2(eps[:, :, ::4, ::4], T(full[f"full_eps_{lab}_sub4"]))
        nrm = float(eps.double().norm())
        if err >= EPS_TOL:
            _, rep = _layer_report(net, sd, cfg, x, labels, cond)
            pytest.fail(f"full model label {lab}: eps rel-L2 {err:.3e}; per-module: {rep}")
        assert abs(nrm / float(full[f"full_eps_{lab}_norm"]) - 1) < 2e-2


def test_eps_batch_vs_oracle_on_gpu():
    """B=3 (odd batch, partial 128-row tiles at the 8x8 level) against the oracle run in fp32 on the GPU."""
    cfg, net, sd = build(common.full_config, 9)
    g = torch.Generator(device=DEV).manual_seed(21)
    x = torch.randn(3, 15, 128, 128, device=DEV, generator=g)
    cond = torch.rand(3, 6, 128, 128, device=DEV, generator=g) * 2 - 1
    labels = torch.full((3,), 500, dtype=torch.long, device=DEV)
    eps = net(x, labels, cond=cond)
    ref, rep = _layer_report(net, sd, cfg, x, labels, cond)
    err = common.rel_l2(eps, ref)
    assert err < EPS_TOL, f"rel-L2 {err:.3e}; per-module: {rep}"
    # batch independence: sample 1 alone gives the same eps (no cross-sample op on the path)
    e1 = net(x[1:2], labels[:1], cond=cond[1:2])
    ind = common.rel_l2(e1, eps[1:2])
    assert ind < 2e-2, f"batch independence {ind:.3e}"  # bf16 rounding noise floor (tile shapes differ with B)


def _tape(seed, n, shape):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(shape, generator=g) for _ in range(n)]


@pytest.mark.parametrize("tag,cfgf,seed,active", [("tiny_act", common.tiny_config, 1, True),
                                                   ("tiny_def", common.tiny_config, 1, False),
                                                   ("gpu64", common.gpu64_config, 4, True)])
def test_samplers_vs_golden(small, tag, cfgf, seed, active):
    from evcdiff import models as M
    cfg, net, sd = build(cfgf, seed, active)
    x_T = T(small[f"{tag}_xT"])
    cond = 2 * T(small[f"{tag}_cond01"]) - 1.0  # float64 like city_sender.py
    tape = _tape(int(small[f"{tag}_ddpm10_noise_seed"]), int(small[f"{tag}_ddpm10_n_noise"]), x_T.shape)
    y = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=10,
                       clip_before=True, verbose=True, log=True, noise=tape)
    assert y.shape == (1,) + tuple(x_T.shape) and y.dtype == torch.float32
    ref = T(small[f"{tag}_ddpm10"])
    # 10 coarse steps (skip 100) are less contractive than the 100-step schedule north_star quotes 1e-2 for; the
    # 100-step bound is tested in test_ddpm100_trajectory
    assert common.rel_l2(y, ref) < 2.5e-2, ("ddpm", common.rel_l2(y, ref))
    fr = lambda z: torch.clamp((z + 1) / 2, 0, 1)
    assert common.psnr(fr(y), fr(ref)) > 35.0
    # DDIM / F-PNDM free-running trajectories are chaotic with untrained weights (a 1e-6 perturbation
    # grows to O(1)), so they are checked by decomposition instead: (i) eps parity of the network (tests above),
    # (ii) sampler arithmetic: our graph-captured sampler against the ORACLE sampler (pinned to the reference on
    # CPU, tests/test_oracle.py) when both drive the same, bit-reproducible evcdiff network.
    model = lambda xx, yy: net(xx, yy, cond=cond)
    sched = (net.betas, net.alphas, net.alphas_prev)
    y = M.FPNDM_sampler(x_T.clone(), net, cond=cond, final_only=True, subsample_steps=10, clip_before=True)
    seen = []
    ref = S.fpndm_sampler(x_T.clone(), model, sched, subsample_steps=10, labels_seen=seen)
    assert np.allclose(np.array(seen), small[f"{tag}_fpndm10_labels"])
    assert common.rel_l2(y[0], ref) < 2e-3, ("fpndm", common.rel_l2(y[0], ref))
    y = M.ddim_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=10,
                       clip_before=True)
    ref = S.ddim_sampler(x_T.clone(), model, sched, subsample_steps=10)
    assert common.rel_l2(y[0], ref) < 2e-3, ("ddim", common.rel_l2(y[0], ref))
    y = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=10,
                       clip_before=True, noise=tape)
    ref = S.ddpm_sampler(x_T.clone(), model, sched, lambda i: tape[i].to(DEV), subsample_steps=10)
    assert common.rel_l2(y[0], ref) < 2e-3, ("ddpm", common.rel_l2(y[0], ref))
    imgs = M.FPNDM_sampler(x_T.clone(), net, cond=cond, final_only=False, subsample_steps=10)
    assert imgs.shape == (10,) + tuple(x_T.shape) and imgs.device.type == "cpu"


def test_teacher_forced_steps_and_graph_equals_eager():
    """Per-step x_t / eps against the oracle fed with OUR previous x_t (teacher forcing), DDPM and DDIM; then the
    captured-graph loop must reproduce the eager loop bit for bit when fed the same generator state."""
    from evcdiff import models as M
    cfg, net, sd = build(common.gpu64_config, 4)
    g = torch.Generator(device=DEV).manual_seed(5)
    B = 2
    x_T = torch.randn(B, 15, 32, 32, device=DEV, generator=g)
    cond = torch.rand(B, 6, 32, 32, device=DEV, generator=g, dtype=torch.float64) * 2 - 1
    sched = S.schedule(cfg, DEV)
    model = lambda x, y: O.ncsnpp_forward(sd, cfg, x, y, cond)
    tape = [t.to(DEV) for t in _tape(77, 19, x_T.shape)]
    imgs = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=False, denoise=True, subsample_steps=20,
                          clip_before=True, noise=tape)
    assert imgs.shape[0] == 21 and imgs.device.type == "cpu"
    trace = []
    S.ddpm_sampler(x_T.clone(), model, sched, lambda i: tape[i], subsample_steps=20, trace=trace)
    worst = max(common.rel_l2(imgs[i].to(DEV), trace[i][0]) for i in range(21))
    assert worst < 2.5e-2, worst
    # graph vs eager with the global generator
    torch.manual_seed(1234)
    torch.cuda.manual_seed_all(1234)
    a = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=True, subsample_steps=20, graph=False)
    torch.manual_seed(1234)
    torch.cuda.manual_seed_all(1234)
    b = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=True, subsample_steps=20, graph=True)
    torch.manual_seed(1234)
    torch.cuda.manual_seed_all(1234)
    c = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=True, subsample_steps=20, graph=True)  # replay
    assert torch.equal(b, c)
    assert torch.equal(a, b)  # eager launches == captured graph, bit for bit (fixed-point integer atomics, fixed split-K order)
    # the same generator draws as the reference loop: randn_like per step after seeding
    torch.manual_seed(1234)
    torch.cuda.manual_seed_all(1234)
    ref = S.ddpm_sampler(x_T.clone(), model, sched, lambda i: torch.randn_like(x_T), subsample_steps=20)
    assert common.rel_l2(b[0], ref) < 2.5e-2, common.rel_l2(b[0], ref)


def test_ddpm100_trajectory():
    """The north_star condition: 100-step DDPM (101 evaluations), per-step x_t rel-L2 <= 1e-2 in bf16 mode against the
    fp32 oracle on the same noise, and final-frame PSNR against a fixed target within 0.05 dB of the oracle's."""
    from evcdiff import models as M
    cfg, net, sd = build(common.gpu64_config, 4)
    g = torch.Generator(device=DEV).manual_seed(15)
    B = 2
    x_T = torch.randn(B, 15, 32, 32, device=DEV, generator=g)
    cond01 = torch.rand(B, 6, 32, 32, device=DEV, generator=g, dtype=torch.float64)
    target = torch.rand(B, 15, 32, 32, device=DEV, generator=g)
    cond = 2 * cond01 - 1
    sched = S.schedule(cfg, DEV)
    model = lambda x, y: O.ncsnpp_forward(sd, cfg, x, y, cond)
    tape = [t.to(DEV) for t in _tape(78, 99, x_T.shape)]
    imgs = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=False, denoise=True, subsample_steps=100,
                          clip_before=True, noise=tape)
    assert imgs.shape[0] == 101
    trace = []
    S.ddpm_sampler(x_T.clone(), model, sched, lambda i: tape[i], subsample_steps=100, trace=trace)
    errs = [common.rel_l2(imgs[i].to(DEV), trace[i][0]) for i in range(101)]
    assert max(errs) < XT_TOL, (max(errs), errs[::10])
    fr = lambda z: torch.clamp((z + 1) / 2, 0, 1)
    p_ours, p_ref = common.psnr(fr(imgs[-1].to(DEV)), target), common.psnr(fr(trace[-1][0]), target)
    assert abs(p_ours - p_ref) < 0.05, (p_ours, p_ref)


def test_ddim_step_sweep_and_pndm_api():
    """BASELINE config 4 (DDIM 10/25/50/100/1000 -> 11/26/51/101/1001 evaluations, one graph per schedule) on the
    gpu64 network at default-like init, and the public pndm functions driven eagerly (reference models/pndm.py)."""
    from evcdiff import models as M
    from evcdiff import ops
    from evcdiff.models import pndm
    cfg, net, sd = build(common.gpu64_config, 4, active=False)
    g = torch.Generator(device=DEV).manual_seed(33)
    x_T = torch.randn(1, 15, 32, 32, device=DEV, generator=g)
    cond = torch.rand(1, 6, 32, 32, device=DEV, generator=g, dtype=torch.float64) * 2 - 1
    model = lambda xx, yy: net(xx, yy, cond=cond)
    sched = (net.betas, net.alphas, net.alphas_prev)
    per_eval = None
    for steps in (10, 25, 50, 100, 1000):
        n0 = ops.launch_count()
        y = M.ddim_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=steps)
        assert y.shape == (1, 1, 15, 32, 32) and torch.isfinite(y).all()
        loop = net.engine(1, DEV)._loop
        launched = max(v for k, v in loop.launches_per_run.items() if k[0] == "ddim" and k[1] == min(steps, 1000))
        evals = steps + 1
        if per_eval is None:
            per_eval = (launched - evals) / evals  # UNet launches + 1 update per evaluation
        assert launched == round(evals * (per_eval + 1)), (steps, launched)
        if steps <= 50:
            ref = S.ddim_sampler(x_T.clone(), model, sched, subsample_steps=steps)
            assert common.rel_l2(y[0], ref) < 2e-3, (steps, common.rel_l2(y[0], ref))
    # pndm API (eager): gen_order_4 for 4 steps = 3 Runge-Kutta steps + 1 Adams-Bashforth step, and gen_order_1
    a_old = net.alphas.flip(0)
    x1, x2, ets1, ets2 = x_T.clone(), x_T.clone(), [], []
    for t, tn in [(0, -1), (100, 0), (200, 100), (300, 200)]:
        tt = torch.full((1,), t, device=DEV).long()
        tnn = torch.full((1,), tn, device=DEV).long()
        x1, ets1 = pndm.gen_order_4(x1, tt, tnn, model, a_old, ets1, clip_before=True)
        x2, ets2 = S.gen_order_4(x2, tt, tnn, model, a_old, ets2, clip_before=True)
    assert len(ets1) == 4 and common.rel_l2(x1, x2) < 2e-3
    tt, tnn = torch.full((1,), 500, device=DEV).long(), torch.full((1,), 400, device=DEV).long()
    y1, _ = pndm.gen_order_1(x_T, tt, tnn, model, a_old, [], clip_before=False)
    y2, _ = S.gen_order_1(x_T, tt, tnn, model, a_old, [], clip_before=False)
    assert common.rel_l2(y1, y2) < 1e-5
