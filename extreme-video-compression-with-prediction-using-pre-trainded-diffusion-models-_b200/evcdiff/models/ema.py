"""EMA shadow weights with the reference's interface (reference models/ema.py:4-47): the sender loads the
checkpoint's EMA dict with `load_state_dict` and copies it into the model with `ema` (city_sender.py:317-322).
Pure parameter bookkeeping (no kernels).  `ema()` copies with `p.copy_()` under no_grad (which bumps the version
counter the engines watch; `p.data.copy_()` would not) and also bumps the model's explicit weights epoch, so the
bf16 operands are repacked before the next sampling call."""
import torch
import torch.nn as nn


def _trainable(module):
    """(name, parameter) pairs of the wrapped model, names without the DataParallel 'module.' prefix."""
    inner = module.module if isinstance(module, nn.DataParallel) else module
    return [(n, p) for n, p in inner.named_parameters() if p.requires_grad]


class EMAHelper(object):
    """shadow <- mu * shadow + (1 - mu) * param after every optimiser step; `ema()` writes the shadow into a model."""

    def __init__(self, mu=0.999):
        self.mu = mu
        self.shadow = {}

    def register(self, module):
        self.shadow.update({n: p.data.clone() for n, p in _trainable(module)})

    @torch.no_grad()
    def update(self, module):
        keep = self.mu
        for n, p in _trainable(module):
            self.shadow[n] = keep * self.shadow[n] + (1.0 - keep) * p.data

    @torch.no_grad()
    def ema(self, module):
        for n, p in _trainable(module):
            p.copy_(self.shadow[n])  # in place on the parameter itself: bumps p._version (p.data.copy_ does not)
        inner = module.module if isinstance(module, nn.DataParallel) else module
        for m in inner.modules():  # engines also compare an explicit epoch (robust against `.data` edits made by callers)
            if hasattr(m, "_weights_epoch"):
                m._weights_epoch += 1

    def ema_copy(self, module):
        wrapped = isinstance(module, nn.DataParallel)
        src = module.module if wrapped else module
        twin = type(src)(src.config).to(src.config.device)
        twin.load_state_dict(src.state_dict())
        twin = nn.DataParallel(twin) if wrapped else twin
        self.ema(twin)
        return twin

    def state_dict(self):
        return self.shadow

    def load_state_dict(self, state_dict):
        self.shadow = state_dict
