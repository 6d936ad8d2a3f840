"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on seeded
inputs.  Run in the build container only (the reference does not travel to the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Weights are not stored: both sides regenerate them with tests/common.seeded_state_dict (CPU generator, seed in
the file); load_state_dict(strict=True) below pins the parameter names/shapes the oracle and the product
claim to be state-dict compatible with (446 entries for configs/mine.yml).
"""
import os
import sys

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("EVC_REF", "/root/reference")
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import common  # noqa: E402
from oracle import ncsnpp as O  # noqa: E402

import models as ref_models  # noqa: E402  (reference)
from models import pndm as ref_pndm  # noqa: E402
from models.better.ncsnpp_more import UNetMore_DDPM  # noqa: E402
from models.better import up_or_down_sampling as ref_updown  # noqa: E402
from models.better.layers import get_timestep_embedding as ref_temb  # noqa: E402

torch.set_num_threads(8)


def build_ref(cfg, seed, active):
    net = UNetMore_DDPM(cfg).eval()
    shapes = O.ncsnpp_param_shapes(cfg)
    sd = common.seeded_state_dict(shapes, seed=seed, active=active)
    full = dict(net.state_dict())
    n_param_keys = len([k for k in full if k not in ("betas", "alphas", "alphas_prev", "unet.sigmas")])
    assert n_param_keys == len(sd), (n_param_keys, len(sd))
    full.update(sd)
    net.load_state_dict(full, strict=True)
    return net, sd


class NoiseTape:
    """Replaces torch.randn_like inside the reference samplers by a seeded CPU tape."""

    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)
        self.tape = []

    def __call__(self, x):
        n = torch.randn(x.shape, generator=self.g, dtype=x.dtype)
        self.tape.append(n)
        return n


def run_samplers(net, cfg, tag, out, B, sub_ddpm, sub_ddim, sub_pndm, seed):
    H = cfg.data.image_size
    g = torch.Generator().manual_seed(seed)
    x_T = torch.randn(B, 15, H, H, generator=g)
    cond01 = torch.rand(B, 6, H, H, generator=g, dtype=torch.float64)
    cond = 2 * cond01 - 1.0  # float64, as city_sender.py feeds it
    out[f"{tag}_xT"] = x_T.numpy()
    out[f"{tag}_cond01"] = cond01.numpy()
    orig = torch.randn_like
    tape = NoiseTape(seed + 1)
    torch.randn_like = tape
    try:
        y = ref_models.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=True,
                                    subsample_steps=sub_ddpm, clip_before=True, verbose=True, log=True)
    finally:
        torch.randn_like = orig
    out[f"{tag}_ddpm{sub_ddpm}"] = y.numpy()
    out[f"{tag}_ddpm{sub_ddpm}_noise_seed"] = np.array(seed + 1)
    out[f"{tag}_ddpm{sub_ddpm}_n_noise"] = np.array(len(tape.tape))
    y = ref_models.ddim_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=sub_ddim,
                                clip_before=True, verbose=True, log=True)
    out[f"{tag}_ddim{sub_ddim}"] = y.numpy()
    labels = []
    fwd = net.forward

    def spy(x, y_, cond=None, cond_mask=None):
        labels.append(float(y_[0]))
        return fwd(x, y_, cond=cond, cond_mask=cond_mask)

    net.forward = spy
    try:
        y = ref_models.FPNDM_sampler(x_T.clone(), net, cond=cond, final_only=True, subsample_steps=sub_pndm,
                                     clip_before=True)
    finally:
        net.forward = fwd
    out[f"{tag}_fpndm{sub_pndm}"] = y.numpy()
    out[f"{tag}_fpndm{sub_pndm}_labels"] = np.array(labels, dtype=np.float64)


def make_unet_plain():
    """models/unet.py UNet_DDPM (BASELINE config 5 variant): 'deep' and 'deeper' at a toy size."""
    from models.unet import UNet_DDPM
    from oracle import unet_plain as U
    out = {}
    for mode in ("deep", "deeper"):
        cfg = common.make_config(ngf=32, image_size=16)
        cfg.mode = mode
        net = UNet_DDPM(cfg).eval()
        shapes = U.unet_param_shapes(cfg)
        sd = common.seeded_state_dict(shapes, seed=21, active=True)
        full = dict(net.state_dict())
        assert [k for k in full if k not in ("betas", "alphas", "alphas_prev")] == list(sd.keys()), mode
        full.update(sd)
        net.load_state_dict(full, strict=True)
        g = torch.Generator().manual_seed(22)
        x = torch.randn(2, 15, 16, 16, generator=g)
        # fp32: unlike NCSNpp (ncsnpp_more.py:293) models/unet.py never casts, a float64 cond crashes its first conv
        cond = torch.rand(2, 6, 16, 16, generator=g) * 2 - 1
        out[f"{mode}_x"], out[f"{mode}_cond"] = x.numpy(), cond.numpy()
        for lab in (0, 990):
            with torch.no_grad():
                out[f"{mode}_eps_{lab}"] = net(x, torch.full((2,), lab, dtype=torch.long), cond=cond).numpy()
        orig = torch.randn_like
        tape = NoiseTape(23)
        torch.randn_like = tape
        try:
            y = ref_models.ddpm_sampler(x.clone(), net, cond=cond, final_only=True, denoise=True, subsample_steps=10,
                                        clip_before=True, verbose=True, log=True)
        finally:
            torch.randn_like = orig
        out[f"{mode}_ddpm10"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "unet_plain.npz"), **out)
    print("wrote unet_plain.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


def main():
    if "--unet-plain-only" in sys.argv:
        return make_unet_plain()
    make_unet_plain()
    out = {}
    # ---- schedule + embedding + FIR (config independent pieces)
    cfg = common.tiny_config()
    net, _ = build_ref(cfg, seed=1, active=True)
    for k in ("betas", "alphas", "alphas_prev"):
        out[f"sched_{k}"] = getattr(net, k).numpy()
    t = torch.tensor([0.0, 10.0, 990.0, 999.0, -0.5, 25.0, -1.0])
    out["temb_in"] = t.numpy()
    out["temb_192"] = ref_temb(t, 192).numpy()
    g = torch.Generator().manual_seed(5)
    xf = torch.randn(2, 5, 8, 8, generator=g)
    out["fir_in"] = xf.numpy()
    out["fir_up"] = ref_updown.upsample_2d(xf, (1, 3, 3, 1), factor=2).numpy()
    out["fir_down"] = ref_updown.downsample_2d(xf, (1, 3, 3, 1), factor=2).numpy()

    # ---- pndm.transfer known answers
    alphas_old = net.alphas.flip(0)
    xt = torch.randn(2, 15, 4, 4, generator=g)
    et = torch.randn(2, 15, 4, 4, generator=g)
    tt = torch.tensor([50.0, 50.0])
    tn = torch.tensor([25.0, 25.0])
    out["transfer_x"], out["transfer_et"] = xt.numpy(), et.numpy()
    out["transfer_out"] = ref_pndm.transfer(xt, tt, tn, et, alphas_old, clip_before=True).numpy()
    out["transfer_out_neg"] = ref_pndm.transfer(xt, torch.tensor([0.0, 0.0]), torch.tensor([-0.5, -0.5]), et,
                                                alphas_old, clip_before=False).numpy()

    # ---- tiny config (ngf 32, 16x16): eps + samplers, active and default-like init
    for active in (True, False):
        tag = "tiny_act" if active else "tiny_def"
        net, _ = build_ref(cfg, seed=1, active=active)
        g = torch.Generator().manual_seed(2)
        x = torch.randn(2, 15, 16, 16, generator=g)
        cond = torch.rand(2, 6, 16, 16, generator=g, dtype=torch.float64) * 2 - 1
        out[f"{tag}_x"], out[f"{tag}_cond"] = x.numpy(), cond.numpy()
        for lab in (0, 500, 990):
            with torch.no_grad():
                out[f"{tag}_eps_{lab}"] = net(x, torch.full((2,), lab, dtype=torch.long), cond=cond).numpy()
        with torch.no_grad():
            out[f"{tag}_eps_m0p5"] = net(x, torch.full((2,), -0.5), cond=cond).numpy()
        run_samplers(net, cfg, tag, out, B=2, sub_ddpm=10, sub_ddim=10, sub_pndm=10, seed=3)

    # ---- gpu64 config (ngf 64, 32x32): the smallest shape the CUDA path runs
    cfg64 = common.gpu64_config()
    net, _ = build_ref(cfg64, seed=4, active=True)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 15, 32, 32, generator=g)
    cond = torch.rand(2, 6, 32, 32, generator=g, dtype=torch.float64) * 2 - 1
    out["gpu64_x"], out["gpu64_cond"] = x.numpy(), cond.numpy()
    for lab in (0, 990):
        with torch.no_grad():
            out[f"gpu64_eps_{lab}"] = net(x, torch.full((2,), lab, dtype=torch.long), cond=cond).numpy()
    run_samplers(net, cfg64, "gpu64", out, B=2, sub_ddpm=10, sub_ddim=10, sub_pndm=10, seed=8)
    np.savez_compressed(os.path.join(HERE, "ncsnpp_small.npz"), **out)
    print("wrote ncsnpp_small.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")

    # ---- full configs/mine.yml model (262 M parameters, 128x128), B=1: eps sub-sampled 4x4 to keep the file small
    out = {}
    cfgF = common.full_config()
    net, sd = build_ref(cfgF, seed=9, active=True)
    out["n_params"] = np.array(sum(v.numel() for v in sd.values()))
    g = torch.Generator().manual_seed(10)
    x = torch.randn(1, 15, 128, 128, generator=g)
    cond = torch.rand(1, 6, 128, 128, generator=g, dtype=torch.float64) * 2 - 1
    out["full_x_seed"] = np.array(10)
    for lab in (0, 990):
        with torch.no_grad():
            e = net(x, torch.full((1,), lab, dtype=torch.long), cond=cond)
        out[f"full_eps_{lab}_sub4"] = e[:, :, ::4, ::4].numpy()
        out[f"full_eps_{lab}_norm"] = np.array(float(e.double().norm()))
    np.savez_compressed(os.path.join(HERE, "ncsnpp_full.npz"), **out)
    print("wrote ncsnpp_full.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
