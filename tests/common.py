"""Shared test helpers: configs, deterministic weights, error metrics.  No reference imports here."""
import argparse
import math

import torch


def ns(**kw):
    return argparse.Namespace(**kw)


def make_config(ngf=192, image_size=128, ch_mult=(1, 1, 2, 3, 4), attn_resolutions=(8, 16, 32), n_head_channels=192,
                num_res_blocks=2, device="cpu"):
    """Namespace with the hot-path keys of configs/mine.yml (SURVEY.md section 5 'Config / flags')."""
    return ns(
        device=device,
        data=ns(dataset="Cityscapes", image_size=image_size, channels=3, logit_transform=False,
                uniform_dequantization=False, gaussian_dequantization=False, rescaled=True, num_frames=5,
                num_frames_cond=2, num_frames_future=0),
        model=ns(depth="deeper", version="DDPM", gamma=False, arch="unetmore", type="v1", time_conditional=True,
                 dropout=0.0, sigma_dist="linear", sigma_begin=0.02, sigma_end=0.0001, num_classes=1000, ema=True,
                 ema_rate=0.999, spec_norm=False, normalization="InstanceNorm++", nonlinearity="swish", ngf=ngf,
                 ch_mult=list(ch_mult), num_res_blocks=num_res_blocks, attn_resolutions=list(attn_resolutions),
                 n_head_channels=n_head_channels, conditional=True, noise_in_cond=False, output_all_frames=False,
                 cond_emb=False, spade=False, spade_dim=128),
        sampling=ns(subsample=100, denoise=True, clip_before=True, init_prev_t=-1.0, final_only=True, step_lr=0.0,
                    n_steps_each=0, num_frames_pred=28, ckpt_id=0),
    )


# tiny = runs in the reference on CPU in seconds; gpu64 = smallest config the CUDA path supports (channels % 64 == 0)
def tiny_config(device="cpu"):
    return make_config(ngf=32, image_size=16, attn_resolutions=(2, 4, 8), n_head_channels=32, device=device)


def gpu64_config(device="cpu"):
    return make_config(ngf=64, image_size=32, attn_resolutions=(2, 4, 8), n_head_channels=64, device=device)


def full_config(device="cpu"):
    return make_config(device=device)


def seeded_state_dict(shapes, seed=0, active=True, device="cpu"):
    """Deterministic fp32 weights for a {key: shape} dict (CPU generator => identical on every machine).

    active=True gives every tensor O(1) effect ("active init", SURVEY.md section 4 trap: the reference's
    zero-init layers make default-init outputs ~1e-5 and parity vacuous)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = {}
    for k, shp in shapes.items():
        if k.endswith("Norm_0.weight") or k.endswith("GroupNorm_0.weight") or k.endswith("normalize.weight") \
                or "normalize0.weight" in k or "normalize1.weight" in k:
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias") or k.endswith(".b"):
            t = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("Dense_0.weight"):
            t = torch.randn(shp, generator=g) * (0.5 / math.sqrt(shp[1]))
        elif k.endswith(".W"):  # NIN (in, out)
            t = torch.randn(shp, generator=g) / math.sqrt(shp[0])
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            t = torch.randn(shp, generator=g) / math.sqrt(fan_in)
        if not active and (k.endswith("Conv_1.weight") or k.endswith("NIN_3.W")):
            t = t * 1e-5
        sd[k] = t.to(device)
    return sd


def rel_l2(a, b):
    a = a.double().flatten()
    b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 10.0 * math.log10(1.0 / max(mse, 1e-30))
