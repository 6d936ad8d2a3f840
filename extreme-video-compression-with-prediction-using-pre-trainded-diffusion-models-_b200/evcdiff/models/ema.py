"""EMA shadow weights with the reference's interface (reference models/ema.py:4-47): the sender loads the
checkpoint's EMA dict with `load_state_dict` and copies it into the model with `ema` (city_sender.py:317-322).
Pure parameter bookkeeping (no kernels); the engine notices the in-place `copy_` through the parameters'
version counters and repacks its bf16 operands before the next sampling call."""
import torch.nn as nn


def _unwrap(module):
    return module.module if isinstance(module, nn.DataParallel) else module


class EMAHelper(object):
    def __init__(self, mu=0.999):
        self.mu = mu
        self.shadow = {}

    def register(self, module):
        for name, param in _unwrap(module).named_parameters():
            if param.requires_grad:
                self.shadow[name] = param.data.clone()

    def update(self, module):
        for name, param in _unwrap(module).named_parameters():
            if param.requires_grad:
                self.shadow[name].data = (1.0 - self.mu) * param.data + self.mu * self.shadow[name].data

    def ema(self, module):
        for name, param in _unwrap(module).named_parameters():
            if param.requires_grad:
                param.data.copy_(self.shadow[name].data)

    def ema_copy(self, module):
        inner = _unwrap(module)
        copy = type(inner)(inner.config).to(inner.config.device)
        copy.load_state_dict(inner.state_dict())
        if isinstance(module, nn.DataParallel):
            copy = nn.DataParallel(copy)
        self.ema(copy)
        return copy

    def state_dict(self):
        return self.shadow

    def load_state_dict(self, state_dict):
        self.shadow = state_dict
