"""Samplers of the conditional video-diffusion hot path, with the reference's names and call signatures
(reference models/__init__.py:17-342).  `scorenet` must be an evcdiff model (UNetMore_DDPM / UNet_DDPM); every
UNet evaluation and every sampler update runs in libevcdiff.so, and by default the whole multi-step loop is
captured once per (sampler, batch, schedule) into a single CUDA graph and replayed.

Behaviour kept from the reference (quirks included):
  * schedule buffers come from the model (`alphas`, `alphas_prev`, `betas`, index 0 = noisiest);
  * sub-sampling recomputes betas as 1 - alpha/alpha_prev; the last step adds no noise; the final denoise
    evaluation uses the label L-1 (an index, not a timestep);
  * FPNDM: ascending indices over the flipped alphas, fractional / negative labels (0, -0.5, -0.5, -1, ...),
    `denoise` ignored, no noise drawn, `subsample_steps=None` is an error;
  * Gaussian noise is drawn from the global CUDA generator with the same calls (one `normal_` per step over
    (B,15,H,W) fp32), so seeding torch the same way gives the same noise sequence as the reference on GPU;
  * `final_only=True` returns `x.unsqueeze(0)`; otherwise the per-step `x.to('cpu')` stack;
  * unknown keyword arguments are swallowed.
Not supported (raise, no fallback): gamma noise, t_min > 0 (noised-start), frac_steps.
Extra optional keywords: `noise` (sequence / tensor of per-step noise, used instead of the generator: parity tests
feed the oracle's noise through it, and the multi-GPU path feeds each rank its slice of the global-batch noise;
with graph=True it is copied into a static tape the captured graph reads), `graph` (False = eager launches,
default True) and `precision` ('bf16' = the model's default | 'fp32': split-bf16 x3 tensor-core arithmetic, per-step
x_t within 1e-3 of the fp32 reference; applies to this call only).
"""
import numpy as np
import torch

from .._lib import EvcError, StepCoef


def get_sigmas(config):
    """Noise schedule (reference models/__init__.py:17-36)."""
    T = getattr(config.model, "num_classes")
    dist = config.model.sigma_dist
    if dist == "geometric":
        return torch.logspace(np.log10(config.model.sigma_begin), np.log10(config.model.sigma_end), T).to(config.device)
    if dist == "linear":
        return torch.linspace(config.model.sigma_begin, config.model.sigma_end, T).to(config.device)
    if dist == "cosine":
        t = torch.linspace(T, 0, T + 1) / T
        s = 0.008
        f = torch.cos((t + s) / (1 + s) * np.pi / 2) ** 2
        return f[:-1] / f[-1]
    raise NotImplementedError("sigma distribution not supported")


from . import pndm  # noqa: E402  (after get_sigmas: the model modules import it from here)
from .loop import SamplerLoop  # noqa: E402


def _net_of(scorenet):
    net = scorenet.module if hasattr(scorenet, "module") else scorenet
    if not hasattr(net, "engine"):
        raise EvcError("scorenet must be an evcdiff model (UNetMore_DDPM / UNet_DDPM); torch modules are not "
                       "evaluated by this package")
    return net


def _subsampled_schedule(net, subsample_steps):
    alphas, alphas_prev, betas = net.alphas, net.alphas_prev, net.betas
    steps = np.arange(len(betas))
    if subsample_steps is not None and subsample_steps < len(alphas):
        skip = len(alphas) // subsample_steps
        steps = torch.tensor(range(0, len(alphas), skip), device=alphas.device)
        alphas = alphas.index_select(0, steps)
        alphas_prev = torch.cat([alphas[1:], torch.tensor([1.0]).to(alphas)])
        betas = 1.0 - torch.div(alphas, alphas_prev)
        steps = steps.cpu().numpy()
    return steps, alphas, alphas_prev, betas


def _unsupported(gamma, t_min, frac_steps=None):
    if gamma:
        raise EvcError("gamma noise is not supported on the B200 sampling path (configs/mine.yml: model.gamma=false)")
    if t_min is not None and t_min > 0:
        raise EvcError("t_min > 0 (noised start from previous frames) is not supported (sampling.init_prev_t=-1)")
    if frac_steps is not None:
        raise EvcError("frac_steps is not supported")


def _ancestral(kind, x_mod, scorenet, cond, final_only, denoise, subsample_steps, clip_before, just_beta, same_noise,
               noise_val, noise, graph, precision=None):
    net = _net_of(scorenet)
    steps, alphas, alphas_prev, betas = _subsampled_schedule(net, subsample_steps)
    L = len(steps)
    # per-step coefficients, evaluated with the same fp32 tensor expressions as the reference
    k0 = (1 / alphas.sqrt())
    k1 = (1 - alphas).sqrt()
    if kind == "ddpm":
        c_x0 = alphas_prev.sqrt() * betas / (1 - alphas)
        c_x = (1 - betas).sqrt() * (1 - alphas_prev) / (1 - alphas)
        c_eps = torch.zeros_like(alphas)
        c_noise = betas.sqrt() if just_beta else ((1 - alphas_prev) / (1 - alphas) * betas).sqrt()
    else:
        c_x0 = alphas_prev.sqrt()
        c_x = torch.zeros_like(alphas)
        c_eps = (1 - alphas_prev).sqrt()
        c_noise = torch.zeros_like(alphas)
    tab = torch.stack([k0, k1, c_x0, c_x, c_eps, c_noise], 1).float().cpu().tolist()
    coefs, labels, draws = [], [], []
    for i in range(L):
        r = tab[i]
        use_noise = kind == "ddpm" and (i + 1 != L)  # the reference draws on every non-final step (:313-326)
        coefs.append(StepCoef(0, int(bool(clip_before)), r[0], r[1], r[2], r[3], r[4], r[5] if use_noise else 0.0))
        labels.append(float(steps[i]))
        draws.append(use_noise)
    if denoise:
        coefs.append(StepCoef(1, 0, 0.0, tab[-1][1], 0.0, 0.0, 0.0, 0.0))
        labels.append(float(L - 1))
        draws.append(False)
    if same_noise and noise_val is None:
        noise_val = x_mod.detach().clone()
    loop = SamplerLoop.get(net, x_mod.shape[0], x_mod.device, precision)  # precision: this call only (None = net.precision)
    key = (kind, L, int(bool(denoise)), int(bool(clip_before)), int(bool(just_beta)), tuple(labels),
           tuple(c.c_noise for c in coefs))
    return loop.run_ancestral(key, x_mod, cond, labels, coefs, final_only=final_only, noise=noise,
                              noise_const=noise_val if same_noise else None, graph=graph, draws=draws)


@torch.no_grad()
def ddpm_sampler(x_mod, scorenet, cond=None, just_beta=False, final_only=False, denoise=True, subsample_steps=None,
                 same_noise=False, noise_val=None, frac_steps=None, verbose=False, log=False, clip_before=True,
                 t_min=-1, gamma=False, noise=None, graph=True, precision=None, **kwargs):
    """Ancestral DDPM sampling (reference models/__init__.py:207-342)."""
    _unsupported(gamma, t_min, frac_steps)
    return _ancestral("ddpm", x_mod, scorenet, cond, final_only, denoise, subsample_steps, clip_before, just_beta,
                      same_noise, noise_val, noise, graph, precision)


@torch.no_grad()
def ddim_sampler(x_mod, scorenet, cond=None, final_only=False, denoise=True, subsample_steps=None, verbose=False,
                 log=True, clip_before=True, t_min=-1, gamma=False, graph=True, precision=None, **kwargs):
    """Deterministic DDIM sampling (reference models/__init__.py:103-204)."""
    _unsupported(gamma, t_min)
    return _ancestral("ddim", x_mod, scorenet, cond, final_only, denoise, subsample_steps, clip_before, False, False,
                      None, None, graph, precision)


@torch.no_grad()
def FPNDM_sampler(x_mod, scorenet, cond=None, final_only=False, denoise=True, subsample_steps=None, verbose=False,
                  log=True, clip_before=True, t_min=-1, gamma=False, graph=True, precision=None, **kwargs):
    """F-PNDM sampling: 3 Runge-Kutta steps then 4th-order Adams-Bashforth (reference models/__init__.py:39-100)."""
    net = _net_of(scorenet)
    alphas = net.alphas
    skip = len(alphas) // subsample_steps  # TypeError on None, like the reference (:62)
    steps = list(range(0, len(alphas), skip))
    steps_next = [-1] + steps[:-1]
    alphas_old = alphas.flip(0).float().cpu()
    loop = SamplerLoop.get(net, x_mod.shape[0], x_mod.device, precision)
    key = ("fpndm", len(steps), skip, int(bool(clip_before)))
    return loop.run_fpndm(key, x_mod, cond, steps, steps_next, alphas_old, clip_before, final_only=final_only,
                          graph=graph)
