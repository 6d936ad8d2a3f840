"""PNDM / F-PNDM step functions with the reference's names and signatures (reference models/pndm.py:3-52),
evaluated by the evc_pndm_update kernel.  `model(x, t)` is any callable returning eps (an evcdiff model's
forward, or functools.partial(net, cond=cond)); tensors are (B,C,H,W) fp32 CUDA tensors.

FPNDM_sampler does not call these one by one: it enqueues the same kernels inside one CUDA graph
(evcdiff/models/loop.py).  They are exported for API completeness and for the per-function parity tests."""
import torch

from .. import ops
from .._lib import EvcError, PndmCoef


def _coef(t, t_next, alphas_cump, n_e, w, w_scale, clip):
    tl, tnl = t.long(), t_next.long()
    if not (bool((tl == tl[0]).all()) and bool((tnl == tnl[0]).all())):
        raise EvcError("per-sample timesteps are not supported: the sampling path uses batch-uniform t")
    ac = alphas_cump.float().cpu()
    at = ac[int(tl[0]) + 1]
    an = ac[int(tnl[0]) + 1]
    c = PndmCoef()
    c.n_e, c.clip = n_e, int(bool(clip))
    for j in range(4):
        c.w[j] = float(w[j]) if j < n_e else 0.0
    c.w_scale = float(torch.tensor(w_scale, dtype=torch.float32))
    c.d = float(an - at)
    c.p = float(1 / (at.sqrt() * (at.sqrt() + an.sqrt())))
    c.q = float(1 / (at.sqrt() * (((1 - an) * at).sqrt() + ((1 - at) * an).sqrt())))
    return c


def _f32(x):
    return x.to(torch.float32).contiguous()


def transfer(x, t, t_next, et, alphas_cump, clip_before=False):
    """x_next = x + (a_next - a) * (x / (sqrt(a)(sqrt(a)+sqrt(a_next))) - et / (sqrt(a)(sqrt((1-a_next)a)+sqrt((1-a)a_next))))."""
    x = _f32(x)
    out = torch.empty_like(x)
    ops.pndm_update(x, [_f32(et)], out, None, None, _coef(t, t_next, alphas_cump, 1, (1.0,), 1.0, clip_before))
    return out


def _combine(es, w, w_scale):
    es = [_f32(e) for e in es]
    et = torch.empty_like(es[0])
    c = PndmCoef()
    c.n_e, c.clip = len(es), 0
    for j in range(4):
        c.w[j] = float(w[j]) if j < len(es) else 0.0
    c.w_scale = float(torch.tensor(w_scale, dtype=torch.float32))
    c.d, c.p, c.q = 0.0, 0.0, 0.0
    scratch = torch.empty_like(es[0])
    ops.pndm_update(es[0], es, scratch, et, None, c)
    return et


def runge_kutta(x, t_list, model, alphas_cump, ets, clip_before=False):
    e_1 = model(x, t_list[0])
    ets.append(e_1)
    x_2 = transfer(x, t_list[0], t_list[1], e_1, alphas_cump, clip_before)
    e_2 = model(x_2, t_list[1])
    x_3 = transfer(x, t_list[0], t_list[1], e_2, alphas_cump, clip_before)
    e_3 = model(x_3, t_list[1])
    x_4 = transfer(x, t_list[0], t_list[2], e_3, alphas_cump, clip_before)
    e_4 = model(x_4, t_list[2])
    et = _combine([e_1, e_2, e_3, e_4], (1.0, 2.0, 2.0, 1.0), 1 / 6)
    return et, ets


def gen_order_1(img, t, t_next, model, alphas_cump, ets, clip_before=False):  # DDIM
    noise = model(img, t)
    ets.append(noise)
    img_next = transfer(img, t, t_next, noise, alphas_cump, clip_before)
    return img_next, ets


def gen_order_4(img, t, t_next, model, alphas_cump, ets, clip_before=False):  # F-PNDM
    t_list = [t, (t + t_next) / 2, t_next]
    if len(ets) > 2:
        noise_ = model(img, t)
        ets.append(noise_)
        noise = _combine([ets[-1], ets[-2], ets[-3], ets[-4]], (55.0, -59.0, 37.0, -9.0), 1 / 24)
    else:
        noise, ets = runge_kutta(img, t_list, model, alphas_cump, ets, clip_before)
    img_next = transfer(img, t, t_next, noise, alphas_cump, clip_before)
    return img_next, ets
