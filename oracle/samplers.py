"""ORACLE (test infrastructure, never imported by the product path).

fp32 torch restatement of the reference samplers, with the Gaussian noise supplied by a callable so a CPU oracle
and the CUDA path can be fed identical noise.  Pinned against the live reference by tests/golden/make_golden.py.

  schedule        models/better/ncsnpp_more.py:735-739 + models/__init__.py:17-36 (linear)
  ddpm_sampler    models/__init__.py:207-342
  ddim_sampler    models/__init__.py:103-204
  fpndm_sampler   models/__init__.py:39-100
  transfer / runge_kutta / gen_order_4 / gen_order_1   models/pndm.py:3-52
Only the configuration the hot path uses is restated (t_min <= 0, gamma False, frac_steps None, same_noise False).
"""
import torch


def schedule(cfg, device="cpu"):
    m = cfg.model
    assert m.sigma_dist == "linear"
    betas = torch.linspace(m.sigma_begin, m.sigma_end, m.num_classes).to(device)
    alphas = torch.cumprod(1 - betas.flip(0), 0).flip(0)
    alphas_prev = torch.cat([alphas[1:], torch.tensor([1.0]).to(alphas)])
    return betas, alphas, alphas_prev


def _subsample(alphas, alphas_prev, betas, subsample_steps):
    steps = torch.arange(len(betas), device=alphas.device)
    if subsample_steps is not None and subsample_steps < len(alphas):
        skip = len(alphas) // subsample_steps
        steps = torch.tensor(range(0, len(alphas), skip), device=alphas.device)
        alphas = alphas.index_select(0, steps)
        alphas_prev = torch.cat([alphas[1:], torch.tensor([1.0]).to(alphas)])
        betas = 1.0 - torch.div(alphas, alphas_prev)
    return steps, alphas, alphas_prev, betas


@torch.no_grad()
def ddpm_sampler(x, model, sched, noise_fn, subsample_steps=None, denoise=True, clip_before=True, trace=None):
    """model(x, labels) -> eps.  noise_fn(i) -> noise for step i (like x).  Returns x (B,C,H,W)."""
    betas, alphas, alphas_prev = sched
    steps, alphas, alphas_prev, betas = _subsample(alphas, alphas_prev, betas, subsample_steps)
    L = len(steps)
    for i, step in enumerate(steps):
        c_beta, c_alpha, c_alpha_prev = betas[i], alphas[i], alphas_prev[i]
        labels = (step * torch.ones(x.shape[0], device=x.device)).long()
        grad = model(x, labels)
        x0 = (1 / c_alpha.sqrt()) * (x - (1 - c_alpha).sqrt() * grad)
        if clip_before:
            x0 = x0.clip_(-1, 1)
        x = (c_alpha_prev.sqrt() * c_beta / (1 - c_alpha)) * x0 + ((1 - c_beta).sqrt() * (1 - c_alpha_prev) / (1 - c_alpha)) * x
        if i + 1 != L:
            x = x + ((1 - c_alpha_prev) / (1 - c_alpha) * c_beta).sqrt() * noise_fn(i)
        if trace is not None:
            trace.append((x.clone(), grad.clone()))
    if denoise:
        last = ((L - 1) * torch.ones(x.shape[0], device=x.device)).long()
        grad = model(x, last)
        x = x - (1 - alphas[-1]).sqrt() * grad
        if trace is not None:
            trace.append((x.clone(), grad.clone()))
    return x


@torch.no_grad()
def ddim_sampler(x, model, sched, subsample_steps=None, denoise=True, clip_before=True, trace=None):
    betas, alphas, alphas_prev = sched
    steps, alphas, alphas_prev, betas = _subsample(alphas, alphas_prev, betas, subsample_steps)
    L = len(steps)
    for i, step in enumerate(steps):
        c_alpha, c_alpha_prev = alphas[i], alphas_prev[i]
        labels = (step * torch.ones(x.shape[0], device=x.device)).long()
        grad = model(x, labels)
        x0 = (1 / c_alpha.sqrt()) * (x - (1 - c_alpha).sqrt() * grad)
        if clip_before:
            x0 = x0.clip_(-1, 1)
        x = c_alpha_prev.sqrt() * x0 + (1 - c_alpha_prev).sqrt() * grad
        if trace is not None:
            trace.append((x.clone(), grad.clone()))
    if denoise:
        last = ((L - 1) * torch.ones(x.shape[0], device=x.device)).long()
        grad = model(x, last)
        x = x - (1 - alphas[-1]).sqrt() * grad
        if trace is not None:
            trace.append((x.clone(), grad.clone()))
    return x


def transfer(x, t, t_next, et, alphas_cump, clip_before=False):
    at = alphas_cump[t.long() + 1].view(-1, 1, 1, 1)
    at_next = alphas_cump[t_next.long() + 1].view(-1, 1, 1, 1)
    x_delta = (at_next - at) * ((1 / (at.sqrt() * (at.sqrt() + at_next.sqrt()))) * x -
                                1 / (at.sqrt() * (((1 - at_next) * at).sqrt() + ((1 - at) * at_next).sqrt())) * et)
    x_next = x + x_delta
    if clip_before:
        x_next = x_next.clip_(-1, 1)
    return x_next


def runge_kutta(x, t_list, model, alphas_cump, ets, clip_before=False):
    e_1 = model(x, t_list[0])
    ets.append(e_1)
    x_2 = transfer(x, t_list[0], t_list[1], e_1, alphas_cump, clip_before)
    e_2 = model(x_2, t_list[1])
    x_3 = transfer(x, t_list[0], t_list[1], e_2, alphas_cump, clip_before)
    e_3 = model(x_3, t_list[1])
    x_4 = transfer(x, t_list[0], t_list[2], e_3, alphas_cump, clip_before)
    e_4 = model(x_4, t_list[2])
    et = (1 / 6) * (e_1 + 2 * e_2 + 2 * e_3 + e_4)
    return et, ets


def gen_order_4(img, t, t_next, model, alphas_cump, ets, clip_before=False):
    t_list = [t, (t + t_next) / 2, t_next]
    if len(ets) > 2:
        noise_ = model(img, t)
        ets.append(noise_)
        noise = (1 / 24) * (55 * ets[-1] - 59 * ets[-2] + 37 * ets[-3] - 9 * ets[-4])
    else:
        noise, ets = runge_kutta(img, t_list, model, alphas_cump, ets, clip_before)
    return transfer(img, t, t_next, noise, alphas_cump, clip_before), ets


def gen_order_1(img, t, t_next, model, alphas_cump, ets, clip_before=False):
    noise = model(img, t)
    ets.append(noise)
    return transfer(img, t, t_next, noise, alphas_cump, clip_before), ets


@torch.no_grad()
def fpndm_sampler(x, model, sched, subsample_steps, clip_before=True, trace=None, labels_seen=None):
    betas, alphas, alphas_prev = sched
    alphas_old = alphas.flip(0)
    skip = len(alphas) // subsample_steps
    steps = list(range(0, len(alphas), skip))
    steps_next = [-1] + steps[:-1]
    steps = torch.tensor(steps, device=alphas.device)
    steps_next = torch.tensor(steps_next, device=alphas.device)
    ets = []

    def wrapped(xx, tt):
        if labels_seen is not None:
            labels_seen.append(float(tt[0]))
        return model(xx, tt)

    for i in range(len(steps)):
        t_ = (steps[i] * torch.ones(x.shape[0], device=x.device)).long()
        t_next = (steps_next[i] * torch.ones(x.shape[0], device=x.device)).long()
        x, ets = gen_order_4(x, t_, t_next, wrapped, alphas_old, ets, clip_before)
        if trace is not None:
            trace.append(x.clone())
    return x


def generate_frames(x_T, cond01, sampler):
    """city_sender.py:326-351 without checkpoint loading: cond in [0,1] -> 2x-1 -> sampler -> (x+1)/2 clamp.
    cond01: (B, 2*3, H, W); sampler(x_T, cond) -> x_0 (B, 15, H, W).  Returns (B, 5, 3, H, W) in [0,1]."""
    cond = 2 * cond01 - 1.0
    x0 = sampler(x_T, cond)
    frames = torch.clamp((x0 + 1.0) / 2.0, 0.0, 1.0)
    B, C, H, W = frames.shape
    return frames.reshape(B, C // 3, 3, H, W)
