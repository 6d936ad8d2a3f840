"""Plain DDPM UNet + DDPM wrapper with the reference's names, constructor, forward signature, buffers and
state-dict keys (reference models/unet.py:49-371; BASELINE config 5), evaluated by libevcdiff.so.

`mode` is read from `config.mode` ('deep' default, 'deeper', 'deepest'), NOT from `model.depth` -- exactly like the
reference (unet.py:184).  UNet_SMLD is out of scope (SMLD samplers are not on the path)."""
import math

import os

import torch
import torch.nn as nn

from .._lib import EvcError
from . import get_sigmas

__all__ = ["UNet_DDPM", "UNet"]


def default_init(module, scale):
    """Xavier-uniform with gain sqrt(scale) (scale 0 -> 1e-10), zero bias (reference unet.py:15-20)."""
    if scale == 0:
        scale = 1e-10
    torch.nn.init.xavier_uniform_(module.weight, math.sqrt(scale))
    torch.nn.init.zeros_(module.bias)


def Normalize(num_channels):
    return nn.GroupNorm(eps=1e-6, num_groups=32, num_channels=num_channels)


class Swish(nn.Module):
    pass  # parameter-free placeholder so temb_dense keeps the reference's indices (0: Linear, 2: Linear)


class Nin(nn.Module):
    """Holds `weights` (out, in) and `bias` like the reference Nin (unet.py:49-59)."""

    def __init__(self, channel_in, channel_out, init_scale=1.0):
        super().__init__()
        self.channel_out = channel_out
        self.weights = nn.Parameter(torch.zeros(channel_out, channel_in), requires_grad=True)
        torch.nn.init.xavier_uniform_(self.weights, math.sqrt(1e-10 if init_scale == 0.0 else init_scale))
        self.bias = nn.Parameter(torch.zeros(channel_out), requires_grad=True)


class ResnetBlock(nn.Module):
    def __init__(self, channel_in, channel_out, tembdim):
        super().__init__()
        self.normalize0 = Normalize(channel_in)
        self.conv0 = nn.Conv2d(channel_in, channel_out, kernel_size=3, padding=1)
        default_init(self.conv0, 1)
        self.dense = nn.Linear(tembdim, channel_out)
        default_init(self.dense, 1)
        self.normalize1 = Normalize(channel_out)
        self.conv1 = nn.Conv2d(channel_out, channel_out, kernel_size=3, padding=1)
        default_init(self.conv1, 0)
        self.nin = Nin(channel_in, channel_out) if channel_in != channel_out else nn.Identity()
        self.channel_in = channel_in


class AttnBlock(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.Q = Nin(channels, channels)
        self.K = Nin(channels, channels)
        self.V = Nin(channels, channels)
        self.OUT = Nin(channels, channels, init_scale=0.0)
        self.normalize = Normalize(channels)
        self.c = channels


class Upsample(nn.Module):
    def __init__(self, channel):
        super().__init__()
        self.conv = nn.Conv2d(channel, channel, kernel_size=3, stride=1, padding=1)
        default_init(self.conv, 1)


def unet_spec(config):
    """Module descriptors of UNet.__init__ (reference unet.py:199-245) in ModuleList order."""
    ch = config.model.ngf
    d = config.data
    mode = getattr(config, "mode", "deep")
    assert mode in ["deep", "deeper", "deepest"]
    mult = {"deepest": (1, 2, 2, 2, 4, 4), "deeper": (1, 2, 2, 4, 4), "deep": (1, 2, 2, 2)}[mode]
    ch_mult = [ch * n for n in mult]
    n_in = d.channels * (d.num_frames + d.num_frames_cond + getattr(d, "num_frames_future", 0))
    down = [dict(kind="conv3", cin=n_in, cout=ch, stride=1)]
    prev_ch, ch_size = ch_mult[0], [ch]
    for i, ich in enumerate(ch_mult):
        for firstarg in [prev_ch, ich]:
            down.append(dict(kind="res", cin=firstarg, cout=ich))
            ch_size.append(ich)
            if i == 1:
                down.append(dict(kind="attn", ch=ich))
        if i != len(ch_mult) - 1:
            down.append(dict(kind="conv3", cin=ich, cout=ich, stride=2))
            ch_size.append(ich)
        prev_ch = ich
    mid = [dict(kind="res", cin=ch_mult[-1], cout=ch_mult[-1]), dict(kind="attn", ch=ch_mult[-1]),
           dict(kind="res", cin=ch_mult[-1], cout=ch_mult[-1])]
    up, prev_ich = [], ch_mult[-1]
    for i, ich in reversed(list(enumerate(ch_mult))):
        for _ in range(3):
            up.append(dict(kind="res", cin=prev_ich + ch_size.pop(), cout=ich))
            if i == 1:
                up.append(dict(kind="attn", ch=ich))
            prev_ich = ich
        if i != 0:
            up.append(dict(kind="upsample", ch=ich))
    if getattr(config.model, "output_all_frames", False):
        raise EvcError("model.output_all_frames is not supported")
    return dict(down=down, mid=mid, up=up, ch=ch, n_out=d.channels * d.num_frames)


class UNet(nn.Module):
    def __init__(self, config):
        super().__init__()
        if not getattr(config.model, "time_conditional", False):
            raise EvcError("models/unet.py path needs model.time_conditional=True")
        if getattr(config.model, "dropout", 0.0) != 0.0:
            raise EvcError("dropout != 0 is not supported (sampling path is eval-only)")
        if config.data.logit_transform or not config.data.rescaled:
            raise EvcError("only data.rescaled=True / logit_transform=False is supported")
        self.config = config
        self.ch = ch = config.model.ngf
        spec = unet_spec(config)

        def make(s):
            if s["kind"] == "res":
                return ResnetBlock(s["cin"], s["cout"], ch * 4)
            if s["kind"] == "attn":
                return AttnBlock(s["ch"])
            if s["kind"] == "upsample":
                return Upsample(s["ch"])
            conv = nn.Conv2d(s["cin"], s["cout"], kernel_size=3, padding=1, stride=s["stride"])
            default_init(conv, 1)
            return conv

        self.downblocks = nn.ModuleList([make(s) for s in spec["down"]])
        self.middleblocks = nn.ModuleList([make(s) for s in spec["mid"]])
        self.upblocks = nn.ModuleList([make(s) for s in spec["up"]])
        self.normalize = Normalize(ch)
        self.out = nn.Conv2d(ch, spec["n_out"], kernel_size=3, stride=1, padding=1)
        default_init(self.out, 0)
        self.temb_dense = nn.Sequential(nn.Linear(ch, ch * 4), Swish(), nn.Linear(ch * 4, ch * 4), Swish())
        default_init(self.temb_dense[0], 1)
        default_init(self.temb_dense[2], 1)
        self._engines = {}
        self._weights_epoch = 0  # bumped by EMAHelper.ema / mark_weights_changed(); part of the engines' staleness key
        # 'bf16' (default, tensor-core speed) or 'fp32' (split-bf16 x3: per-step x_t within 1e-3 of the fp32 reference)
        self.precision = os.environ.get("EVC_PRECISION", getattr(config, "precision", "bf16"))

    def _weights_version(self):
        return (self._weights_epoch, sum(p._version for p in self.parameters()),
                sum(p.data_ptr() % 65521 for p in self.parameters()))

    def mark_weights_changed(self):
        """Call after editing parameters in a way autograd's version counters do not see (e.g. `p.data.copy_()`)."""
        self._weights_epoch += 1

    def engine(self, B, device=None, precision=None):
        """The launch plan for batch size B; `precision` overrides self.precision for this lookup only."""
        from ..engine_unet import PlainUNetEngine
        device = torch.device(device) if device is not None else next(self.parameters()).device
        if device.type != "cuda":
            raise EvcError("evcdiff runs on CUDA devices only; move the model with .to('cuda')")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        precision = precision or self.precision
        key, ver = (B, str(device), precision), self._weights_version()
        hit = self._engines.get(key)
        if hit is None or hit[0] != ver:
            self._engines.pop(key, None)
            with torch.cuda.device(device):
                hit = (ver, PlainUNetEngine(self, B, device, precision))
            self._engines[key] = hit
        return hit[1]

    def forward(self, x, y=None, cond=None):
        if y is None:
            raise EvcError("unconditional (y=None) evaluation is not supported")
        lab = y.float()
        v = float(lab[0])
        if not bool((lab == v).all()):
            raise EvcError("per-sample labels are not supported: the sampling path uses batch-uniform labels")
        eng = self.engine(x.shape[0], x.device)
        with torch.cuda.device(eng.device):
            eng.set_labels([v])
            eng.load_input(x, cond)
            return eng.forward(0).clone()


class UNet_DDPM(nn.Module):
    """DDPM wrapper with the schedule buffers (reference unet.py:327-371)."""

    def __init__(self, config):
        super().__init__()
        self.version = getattr(config.model, "version", "DDPM").upper()
        assert self.version in ("DDPM", "DDIM", "FPNDM"), f"models/unet : version is not DDPM or DDIM! Given: {self.version}"
        for key in ("gamma", "noise_in_cond"):
            if getattr(config.model, key, False):
                raise EvcError(f"model.{key}=True is not supported on the B200 sampling path")
        if getattr(config.model, "sigma_dist", "linear") != "linear":
            raise EvcError("only the linear schedule is supported")
        self.config = config
        self.unet = UNet(config)
        self.schedule = "linear"
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

    def forward(self, x, y, cond=None, labels=None, cond_mask=None):
        return self.unet(x, y, cond)
