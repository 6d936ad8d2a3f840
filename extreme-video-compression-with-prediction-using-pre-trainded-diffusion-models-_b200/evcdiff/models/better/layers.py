"""Parameter containers and initialisers with the reference's names (models/better/layers.py).

These modules only HOLD parameters under the reference's state-dict keys; the forward math runs in
libevcdiff.so through evcdiff.engine, so none of them defines a forward()."""
import math

import numpy as np
import torch
import torch.nn as nn


def variance_scaling(scale, mode="fan_avg", distribution="uniform", in_axis=1, out_axis=0):
    """JAX-style variance scaling (reference layers.py:43-73)."""
    def init(shape, dtype=torch.float32, device="cpu"):
        rf = np.prod(shape) / shape[in_axis] / shape[out_axis]
        fan_in, fan_out = shape[in_axis] * rf, shape[out_axis] * rf
        denom = {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2}[mode]
        var = scale / denom
        if distribution == "normal":
            return torch.randn(*shape, dtype=dtype, device=device) * np.sqrt(var)
        return (torch.rand(*shape, dtype=dtype, device=device) * 2.0 - 1.0) * np.sqrt(3 * var)
    return init


def default_init(scale=1.0):
    """DDPM initialisation (reference layers.py:77-80): scale 0 means 1e-10."""
    return variance_scaling(1e-10 if scale == 0 else scale, "fan_avg", "uniform")


def ddpm_conv3x3(cin, cout, init_scale=1.0):
    conv = nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, bias=True)
    conv.weight.data = default_init(init_scale)(conv.weight.data.shape)
    nn.init.zeros_(conv.bias)
    return conv


def ddpm_conv1x1(cin, cout, init_scale=1.0):
    conv = nn.Conv2d(cin, cout, kernel_size=1, stride=1, padding=0, bias=True)
    conv.weight.data = default_init(init_scale)(conv.weight.data.shape)
    nn.init.zeros_(conv.bias)
    return conv


class NIN(nn.Module):
    """Holds W (in, out) and b like the reference NIN (layers.py:535-540)."""

    def __init__(self, in_dim, num_units, init_scale=0.1):
        super().__init__()
        self.W = nn.Parameter(default_init(scale=init_scale)((in_dim, num_units)), requires_grad=True)
        self.b = nn.Parameter(torch.zeros(num_units), requires_grad=True)


def get_timestep_embedding(timesteps, embedding_dim, max_positions=10000):
    """Sinusoidal embedding (reference layers.py:504-518), computed by the evc_timestep_embedding kernel."""
    from ... import ops
    assert timesteps.dim() == 1
    if embedding_dim % 2 == 1:
        raise NotImplementedError("odd embedding_dim")
    half = embedding_dim // 2
    e = math.log(max_positions) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32, device=timesteps.device) * -e)
    out = torch.empty((timesteps.shape[0], embedding_dim), dtype=torch.float32, device=timesteps.device)
    ops.timestep_embedding(timesteps.float().contiguous(), freqs, embedding_dim, out)
    return out
