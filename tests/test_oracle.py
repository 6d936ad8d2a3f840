"""CPU tests: the oracle (oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  fp32 vs fp32 on CPU: tolerance rel-L2 1e-5 (op order differs slightly:
einsum vs bmm, functional group_norm), trajectories 1e-4 (error growth over 11-29 network evaluations)."""
import os

import numpy as np
import pytest
import torch

import common
from oracle import ncsnpp as O
from oracle import samplers as S

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def small():
    return dict(np.load(os.path.join(G, "ncsnpp_small.npz")))


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_schedule(small):
    cfg = common.tiny_config()
    betas, alphas, alphas_prev = S.schedule(cfg)
    assert torch.equal(betas, T(small["sched_betas"]))
    assert torch.equal(alphas, T(small["sched_alphas"]))
    assert torch.equal(alphas_prev, T(small["sched_alphas_prev"]))
    assert abs(float(alphas[0]) - 4.04e-5) < 1e-6 and abs(float(alphas[-1]) - 0.9999) < 1e-6


def test_timestep_embedding(small):
    e = O.timestep_embedding(T(small["temb_in"]), 192)
    assert common.rel_l2(e, T(small["temb_192"])) < 1e-7


def test_fir(small):
    x = T(small["fir_in"])
    assert common.rel_l2(O.fir_up2(x), T(small["fir_up"])) < 1e-6
    assert common.rel_l2(O.fir_down2(x), T(small["fir_down"])) < 1e-6


def test_fir_closed_form(small):
    """The separable closed form the CUDA kernel implements equals the upfirdn2d definition."""
    x = T(small["fir_in"])
    B, C, H, W = x.shape

    def up1d(t, dim):
        t = t.movedim(dim, -1)
        n = t.shape[-1]
        z = torch.zeros_like(t[..., :1])
        prev = torch.cat([z, t[..., :-1]], -1)
        nxt = torch.cat([t[..., 1:], z], -1)
        even = (prev + 3 * t) / 4
        odd = (3 * t + nxt) / 4
        o = torch.stack([even, odd], -1).reshape(*t.shape[:-1], 2 * n)
        return o.movedim(-1, dim)

    def down1d(t, dim):
        t = t.movedim(dim, -1)
        p = torch.nn.functional.pad(t, (1, 1))
        o = (p[..., 0:-3:2] + 3 * p[..., 1:-2:2] + 3 * p[..., 2:-1:2] + p[..., 3::2]) / 8
        return o.movedim(-1, dim)

    assert common.rel_l2(up1d(up1d(x, 2), 3), T(small["fir_up"])) < 1e-6
    assert common.rel_l2(down1d(down1d(x, 2), 3), T(small["fir_down"])) < 1e-6


def test_transfer(small):
    cfg = common.tiny_config()
    _, alphas, _ = S.schedule(cfg)
    a_old = alphas.flip(0)
    x, et = T(small["transfer_x"]), T(small["transfer_et"])
    y = S.transfer(x, torch.tensor([50.0, 50.0]), torch.tensor([25.0, 25.0]), et, a_old, clip_before=True)
    assert common.rel_l2(y, T(small["transfer_out"])) < 1e-7
    y = S.transfer(x, torch.tensor([0.0, 0.0]), torch.tensor([-0.5, -0.5]), et, a_old, clip_before=False)
    assert common.rel_l2(y, T(small["transfer_out_neg"])) < 1e-7


@pytest.mark.parametrize("tag,active,seed", [("tiny_act", True, 1), ("tiny_def", False, 1)])
def test_eps_tiny(small, tag, active, seed):
    cfg = common.tiny_config()
    sd = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=seed, active=active)
    x, cond = T(small[f"{tag}_x"]), T(small[f"{tag}_cond"])
    for lab in (0, 500, 990):
        e = O.ncsnpp_forward(sd, cfg, x, torch.full((2,), lab, dtype=torch.long), cond)
        assert common.rel_l2(e, T(small[f"{tag}_eps_{lab}"])) < 1e-5
    e = O.ncsnpp_forward(sd, cfg, x, torch.full((2,), -0.5), cond)
    assert common.rel_l2(e, T(small[f"{tag}_eps_m0p5"])) < 1e-5


def _tape(seed, n, shape):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(shape, generator=g) for _ in range(n)]


@pytest.mark.parametrize("tag,cfgf,seed,active", [("tiny_act", common.tiny_config, 1, True),
                                                   ("tiny_def", common.tiny_config, 1, False),
                                                   ("gpu64", common.gpu64_config, 4, True)])
def test_samplers(small, tag, cfgf, seed, active):
    cfg = cfgf()
    sd = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=seed, active=active)
    sched = S.schedule(cfg)
    x_T = T(small[f"{tag}_xT"])
    cond = 2 * T(small[f"{tag}_cond01"]) - 1.0
    model = lambda x, y: O.ncsnpp_forward(sd, cfg, x, y, cond)
    n_noise = int(small[f"{tag}_ddpm10_n_noise"])
    assert n_noise == 9  # last step adds no noise (models/__init__.py:313-315)
    tape = _tape(int(small[f"{tag}_ddpm10_noise_seed"]), n_noise, x_T.shape)
    y = S.ddpm_sampler(x_T.clone(), model, sched, lambda i: tape[i], subsample_steps=10)
    assert common.rel_l2(y.unsqueeze(0), T(small[f"{tag}_ddpm10"])) < 1e-4
    y = S.ddim_sampler(x_T.clone(), model, sched, subsample_steps=10)
    # untrained-weight DDIM trajectories are chaotic: loose bound for the active init
    assert common.rel_l2(y.unsqueeze(0), T(small[f"{tag}_ddim10"])) < (5e-3 if active else 1e-4)
    seen = []
    y = S.fpndm_sampler(x_T.clone(), model, sched, subsample_steps=10, labels_seen=seen)
    assert common.rel_l2(y.unsqueeze(0), T(small[f"{tag}_fpndm10"])) < 1e-4
    assert np.allclose(np.array(seen), small[f"{tag}_fpndm10_labels"])
    assert len(seen) == 3 * 4 + 7  # 3 Runge-Kutta steps + 7 multistep evaluations


def test_eps_gpu64(small):
    cfg = common.gpu64_config()
    sd = common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=4, active=True)
    x, cond = T(small["gpu64_x"]), T(small["gpu64_cond"])
    for lab in (0, 990):
        e = O.ncsnpp_forward(sd, cfg, x, torch.full((2,), lab, dtype=torch.long), cond)
        assert common.rel_l2(e, T(small[f"gpu64_eps_{lab}"])) < 1e-5


def test_param_inventory_full():
    """configs/mine.yml model: 262,133,775 parameters in 442 learnable tensors (+4 schedule buffers = 446)."""
    cfg = common.full_config()
    shapes = O.ncsnpp_param_shapes(cfg)
    n = sum(int(np.prod(s)) for s in shapes.values())
    assert n == 262133775
    assert len(shapes) == 442
    assert shapes["unet.all_modules.3.actnorm0.Dense_0.weight"] == (384, 768)
    full = dict(np.load(os.path.join(G, "ncsnpp_full.npz")))
    assert int(full["n_params"]) == n


@pytest.mark.parametrize("mode", ["deep", "deeper"])
def test_unet_plain(mode):
    """models/unet.py variant (BASELINE config 5): eps and a DDPM-10 trajectory against the reference goldens."""
    from oracle import unet_plain as U
    g = dict(np.load(os.path.join(G, "unet_plain.npz")))
    cfg = common.make_config(ngf=32, image_size=16)
    cfg.mode = mode
    sd = common.seeded_state_dict(U.unet_param_shapes(cfg), seed=21, active=True)
    x, cond = T(g[f"{mode}_x"]), T(g[f"{mode}_cond"])
    for lab in (0, 990):
        e = U.unet_forward(sd, cfg, x, torch.full((2,), lab, dtype=torch.long), cond)
        assert common.rel_l2(e, T(g[f"{mode}_eps_{lab}"])) < 1e-5
    tape = _tape(23, 9, x.shape)
    model = lambda xx, yy: U.unet_forward(sd, cfg, xx, yy, cond)
    y = S.ddpm_sampler(x.clone(), model, S.schedule(cfg), lambda i: tape[i], subsample_steps=10)
    assert common.rel_l2(y.unsqueeze(0), T(g[f"{mode}_ddpm10"])) < 1e-4
    full = common.full_config()
    full.mode = mode
    shapes = U.unet_param_shapes(full)
    n = sum(int(np.prod(s)) for s in shapes.values())
    # ngf=192: 'deep' 80.4 M parameters in 328 tensors (+3 buffers = 331), 'deeper' 240.9 M
    assert (n, len(shapes)) == ((80_434_575, 328) if mode == "deep" else (240_926_991, 390))


def test_sender_restatement_invariants():
    """oracle/sender.py (restatement of city_sender.py:353-437, 519-550) on hand-made cases: accept-all, accept-none and
    a prefix, with the PSNR of city_sender.py:255-258."""
    from oracle import sender as RS
    T, H = 12, 8
    g = torch.Generator().manual_seed(3)
    x_gt = torch.rand(T, 3, H, H, generator=g, dtype=torch.float64)
    perfect = lambda frames: torch.stack([frames[:, 3:6]] * 5, 1)  # repeats the last conditioning frame
    # threshold -inf: everything accepted, 2 keyframes then 5 per cycle
    ge, d, n = RS.encode_video(x_gt, perfect, -1e30, total=T)
    assert d.tolist() == [1, 1] + [0] * 10 and n == 2 and torch.equal(ge[:2], x_gt[:2])
    # threshold +inf: nothing accepted, all keyframes, one (futile) sampling cycle per pair
    ge, d, n = RS.encode_video(x_gt, perfect, 1e30, total=T)
    assert d.tolist() == [1] * T and n == 5 and torch.equal(ge, x_gt)
    # a video whose frames 2..4 equal frame 1 and then jump: the prefix stops at the jump
    x2 = x_gt.clone()
    x2[2:5] = x2[1] + 1e-4
    ge, d, n = RS.encode_video(x2, perfect, 60.0, total=T)
    assert d[:7].tolist() == [1, 1, 0, 0, 0, 1, 1]
    a, b = np.zeros((3, 4, 4)), np.full((3, 4, 4), 0.1)
    assert abs(RS.cal_psnr(a, b) - 20.0) < 1e-9


_LIVE = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1])          # tests/golden
import make_golden as M                   # imports the UNMODIFIED reference from EVC_REF (default /root/reference)
import common
out = {}
cfg = common.tiny_config()
net, _ = M.build_ref(cfg, seed=1, active=True)
for k in ("betas", "alphas", "alphas_prev"):
    out["sched_" + k] = getattr(net, k).numpy()
t = torch.tensor([0.0, 10.0, 990.0, 999.0, -0.5, 25.0, -1.0])
out["temb_192"] = M.ref_temb(t, 192).numpy()
g = torch.Generator().manual_seed(5)
xf = torch.randn(2, 5, 8, 8, generator=g)
out["fir_up"] = M.ref_updown.upsample_2d(xf, (1, 3, 3, 1), factor=2).numpy()
out["fir_down"] = M.ref_updown.downsample_2d(xf, (1, 3, 3, 1), factor=2).numpy()
alphas_old = net.alphas.flip(0)
xt = torch.randn(2, 15, 4, 4, generator=g)
et = torch.randn(2, 15, 4, 4, generator=g)
out["transfer_out"] = M.ref_pndm.transfer(xt, torch.tensor([50.0, 50.0]), torch.tensor([25.0, 25.0]), et, alphas_old,
                                          clip_before=True).numpy()
g = torch.Generator().manual_seed(2)
x = torch.randn(2, 15, 16, 16, generator=g)
cond = torch.rand(2, 6, 16, 16, generator=g, dtype=torch.float64) * 2 - 1
with torch.no_grad():
    out["tiny_act_eps_500"] = net(x, torch.full((2,), 500, dtype=torch.long), cond=cond).numpy()
    out["tiny_act_eps_m0p5"] = net(x, torch.full((2,), -0.5), cond=cond).numpy()
np.savez(sys.argv[2], **out)
"""


@pytest.mark.skipif(not os.path.isdir(os.environ.get("EVC_REF", "/root/reference")),
                    reason="the reference tree only exists in the build container")
def test_goldens_reproduce_from_the_live_reference(small, tmp_path):
    """The committed fixtures ARE the reference's outputs: a subset (schedule, time embedding, FIR up / down, pndm.transfer,
    eps at an integer and at the fractional F-PNDM label) is regenerated here from the unmodified reference, in a separate
    process (its top-level package is called `models`, like ours), and compared with tests/golden/ncsnpp_small.npz."""
    import subprocess
    import sys
    dst = str(tmp_path / "live.npz")
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, "-c", _LIVE, G, dst], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    live = dict(np.load(dst))
    for k, v in live.items():
        ref = small[k]
        assert v.shape == ref.shape and v.dtype == ref.dtype, k
        if k.startswith("sched_"):
            assert np.array_equal(v, ref), k
        else:  # same torch build and thread count as the generator run: equal up to reduction-order noise
            assert common.rel_l2(T(v), T(ref)) < 1e-6, (k, common.rel_l2(T(v), T(ref)))
