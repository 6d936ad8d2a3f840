"""NCSN++ conditional UNet + DDPM wrapper with the reference's names, constructor, forward signature, buffers and
state-dict keys (reference models/better/ncsnpp_more.py:32-392, 721-770), evaluated by libevcdiff.so.

Only the configuration configs/mine.yml instantiates is supported: arch 'unetmore',
positional embedding, BigGAN residual blocks with FIR resampling, GroupNorm/AdaGN, no SPADE, no 3-D variants,
no cond_emb, no noise_in_cond, no gamma.  Anything else raises -- there is no fallback path.
"""
import os

import torch
import torch.nn as nn

from ..._lib import EvcError
from .. import get_sigmas
from . import layers
from .layers import NIN, ddpm_conv1x1, ddpm_conv3x3, default_init

conv3x3 = ddpm_conv3x3
conv1x1 = ddpm_conv1x1


def gn_groups(ch):
    g = min(ch // 4, 32)
    while ch % g != 0:
        g -= 1
    return g


def ncsnpp_spec(config):
    """Module list of NCSNpp in all_modules order (reference ncsnpp_more.py:88-249)."""
    m, d = config.model, config.data
    nf, ch_mult, nres = m.ngf, list(m.ch_mult), m.num_res_blocks
    attn_res = list(m.attn_resolutions)
    nlev = len(ch_mult)
    res = [d.image_size // (2 ** i) for i in range(nlev)]
    n_frames = d.num_frames + d.num_frames_cond + getattr(d, "num_frames_future", 0)
    spec = [dict(kind="linear", cin=nf, cout=nf * 4), dict(kind="linear", cin=nf * 4, cout=nf * 4),
            dict(kind="conv3", cin=d.channels * n_frames, cout=nf, init_scale=1.0)]
    hs_c, in_ch = [nf], nf
    for lvl in range(nlev):
        for _ in range(nres):
            out_ch = nf * ch_mult[lvl]
            spec.append(dict(kind="res", cin=in_ch, cout=out_ch, up=False, down=False))
            in_ch = out_ch
            if res[lvl] in attn_res:
                spec.append(dict(kind="attn", ch=in_ch))
            hs_c.append(in_ch)
        if lvl != nlev - 1:
            spec.append(dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=True))
            hs_c.append(in_ch)
    in_ch = hs_c[-1]
    spec += [dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=False), dict(kind="attn", ch=in_ch),
             dict(kind="res", cin=in_ch, cout=in_ch, up=False, down=False)]
    for lvl in reversed(range(nlev)):
        for _ in range(nres + 1):
            out_ch = nf * ch_mult[lvl]
            spec.append(dict(kind="res", cin=in_ch + hs_c.pop(), cout=out_ch, up=False, down=False))
            in_ch = out_ch
        if res[lvl] in attn_res:
            spec.append(dict(kind="attn", ch=in_ch))
        if lvl != 0:
            spec.append(dict(kind="res", cin=in_ch, cout=in_ch, up=True, down=False))
    assert not hs_c
    spec.append(dict(kind="actnorm_final", ch=in_ch))
    spec.append(dict(kind="conv3", cin=in_ch, cout=d.channels * d.num_frames, init_scale=0.0))
    return spec


class get_act_norm(nn.Module):
    """Parameter holder of the reference get_act_norm (layerspp.py:486-520): AdaGN has Dense_0 only (no GN
    affine), the plain variant has an affine GroupNorm `Norm_0`."""

    def __init__(self, ch, emb_dim=None):
        super().__init__()
        if emb_dim is not None:
            self.Dense_0 = nn.Linear(emb_dim, 2 * ch)
            self.Dense_0.weight.data = default_init()(self.Dense_0.weight.shape)
            nn.init.zeros_(self.Dense_0.bias)
            self.Norm_0 = nn.GroupNorm(gn_groups(ch), ch, eps=1e-5, affine=False)
        else:
            self.Norm_0 = nn.GroupNorm(gn_groups(ch), ch, eps=1e-5, affine=True)


class ResnetBlockBigGANppGN(nn.Module):
    def __init__(self, in_ch, out_ch, temb_dim, up=False, down=False, init_scale=0.0):
        super().__init__()
        self.actnorm0 = get_act_norm(in_ch, temb_dim)
        self.Conv_0 = conv3x3(in_ch, out_ch)
        self.actnorm1 = get_act_norm(out_ch, temb_dim)
        self.Conv_1 = conv3x3(out_ch, out_ch, init_scale=init_scale)
        if in_ch != out_ch or up or down:
            self.Conv_2 = conv1x1(in_ch, out_ch)
        self.up, self.down, self.in_ch, self.out_ch = up, down, in_ch, out_ch


class AttnBlockpp(nn.Module):
    def __init__(self, channels, init_scale=0.0, n_head_channels=-1):
        super().__init__()
        self.GroupNorm_0 = nn.GroupNorm(gn_groups(channels), channels, eps=1e-6)
        self.NIN_0 = NIN(channels, channels)
        self.NIN_1 = NIN(channels, channels)
        self.NIN_2 = NIN(channels, channels)
        self.NIN_3 = NIN(channels, channels, init_scale=init_scale)
        if n_head_channels == -1 or channels < n_head_channels:
            self.n_heads = 1
        else:
            assert channels % n_head_channels == 0
            self.n_heads = channels // n_head_channels


def _check_supported(config):
    m = config.model
    bad = []
    if getattr(m, "arch", "unetmore") != "unetmore":
        bad.append(f"model.arch={m.arch}")
    for key in ("spade", "cond_emb", "noise_in_cond", "gamma", "output_all_frames"):
        if getattr(m, key, False):
            bad.append(f"model.{key}=True")
    if not getattr(m, "time_conditional", True):
        bad.append("model.time_conditional=False")
    if getattr(m, "sigma_dist", "linear") != "linear":
        bad.append(f"model.sigma_dist={m.sigma_dist}")
    if getattr(m, "dropout", 0.0) != 0.0:
        bad.append("model.dropout != 0 (sampling path is eval-only)")
    if bad:
        raise EvcError("unsupported configuration for the B200 sampling path: " + ", ".join(bad))


class NCSNpp(nn.Module):
    """NCSN++ model: same parameter names as the reference (`all_modules.{i}.*`)."""

    def __init__(self, config):
        super().__init__()
        _check_supported(config)
        self.config = config
        self.register_buffer("sigmas", get_sigmas(config))
        self.nf = config.model.ngf
        self.embedding_type = "positional"
        temb_dim = self.nf * 4
        head = getattr(config.model, "n_head_channels", -1)
        mods = []
        for s in ncsnpp_spec(config):
            k = s["kind"]
            if k == "linear":
                lin = nn.Linear(s["cin"], s["cout"])
                lin.weight.data = default_init()(lin.weight.shape)
                nn.init.zeros_(lin.bias)
                mods.append(lin)
            elif k == "conv3":
                mods.append(conv3x3(s["cin"], s["cout"], init_scale=s["init_scale"]))
            elif k == "res":
                mods.append(ResnetBlockBigGANppGN(s["cin"], s["cout"], temb_dim, up=s["up"], down=s["down"]))
            elif k == "attn":
                mods.append(AttnBlockpp(s["ch"], n_head_channels=head))
            elif k == "actnorm_final":
                mods.append(get_act_norm(s["ch"], None))
        self.all_modules = nn.ModuleList(mods)
        self._engines = {}
        self._weights_epoch = 0  # bumped by EMAHelper.ema / mark_weights_changed(); part of the engines' staleness key
        # 'bf16' (default, tensor-core speed) or 'fp32' (split-bf16 x3: per-step x_t within 1e-3 of the fp32 reference)
        self.precision = os.environ.get("EVC_PRECISION", getattr(config, "precision", "bf16"))

    # -- engine management ------------------------------------------------------------------------
    def _weights_version(self):
        return (self._weights_epoch, sum(p._version for p in self.parameters()),
                sum(p.data_ptr() % 65521 for p in self.parameters()))

    def mark_weights_changed(self):
        """Call after editing parameters in a way autograd's version counters do not see (e.g. `p.data.copy_()`)."""
        self._weights_epoch += 1

    def engine(self, B, device=None, precision=None):
        """The launch plan for batch size B (built on first use; rebuilt when the parameters changed).  `precision`
        overrides self.precision for this lookup only."""
        from ...engine import NCSNppEngine
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise EvcError("evcdiff runs on CUDA devices only; move the model with .to('cuda')")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        precision = precision or self.precision
        key = (B, str(device), precision)
        ver = self._weights_version()
        hit = self._engines.get(key)
        if hit is None or hit[0] != ver:
            self._engines.pop(key, None)
            with torch.cuda.device(device):  # plans, TMA maps and the SM count belong to the engine's device
                hit = (ver, NCSNppEngine(self, B, device, precision))
            self._engines[key] = hit
        return hit[1]

    def forward(self, x, time_cond, cond=None, cond_mask=None):
        """eps = NCSNpp(x, labels, cond): one UNet evaluation, (B,15,H,W) fp32 in and out."""
        if cond_mask is not None:
            raise EvcError("cond_mask is not supported (model.cond_emb=False on this path)")
        lab = time_cond.float()
        v = float(lab[0])
        if not bool((lab == v).all()):
            raise EvcError("per-sample labels are not supported: the sampling path uses batch-uniform labels")
        eng = self.engine(x.shape[0], x.device)
        with torch.cuda.device(eng.device):  # kernels are enqueued on the current stream of the engine's device
            eng.set_labels([v])
            eng.load_input(x, cond)
            return eng.forward(0).clone()


class UNetMore_DDPM(nn.Module):
    """DDPM wrapper holding the schedule buffers (reference ncsnpp_more.py:721-770)."""

    def __init__(self, config):
        super().__init__()
        self.version = getattr(config.model, "version", "DDPM").upper()
        assert self.version in ("DDPM", "DDIM", "FPNDM"), f"models/unet : version is not DDPM or DDIM! Given: {self.version}"
        self.config = config
        self.unet = NCSNpp(config)
        self.schedule = getattr(config.model, "sigma_dist", "linear")
        self.register_buffer("betas", get_sigmas(config))
        self.register_buffer("alphas", torch.cumprod(1 - self.betas.flip(0), 0).flip(0))
        self.register_buffer("alphas_prev", torch.cat([self.alphas[1:], torch.tensor([1.0]).to(self.alphas)]))
        self.gamma = False
        self.noise_in_cond = False

    @property
    def precision(self):
        """'bf16' (default) or 'fp32' (split-bf16 x3 tensor-core arithmetic, fp32-tolerance mode)."""
        return self.unet.precision

    @precision.setter
    def precision(self, value):
        self.unet.precision = value

    def engine(self, B, device=None, precision=None):
        return self.unet.engine(B, device, precision)

    def forward(self, x, y, cond=None, cond_mask=None):
        return self.unet(x, y, cond, cond_mask=cond_mask)
