"""CPU test of the HOST side of the sampling loop (evcdiff/models/__init__.py + models/loop.py): label sequence, coefficient
tables, noise draws, denoise step, the F-PNDM step plan with its eps ring and Runge-Kutta / Adams-Bashforth weights.

The real SamplerLoop and the real sampler entry points run here; what is replaced is only what needs a GPU: the UNet engine
(a fake with a smooth closed-form eps) and the two update kernels (torch restatements of `step_value` / `pndm_value` in
csrc/sampler.cu, same operation order).  The result must equal the oracle samplers (oracle/samplers.py, pinned to the reference
by the goldens) driven by the same fake eps.  The kernels themselves are compared with the oracle in tests/test_ncsnpp_gpu.py."""
import contextlib

import pytest
import torch

import common
from oracle import samplers as S


def _fake_eps(x, label, cond):
    lab = torch.as_tensor(label, dtype=torch.float32)
    return 0.6 * torch.tanh(0.7 * x + 0.05 * torch.sin(0.01 * lab)) + 0.2 * cond.float().mean(dim=1, keepdim=True)


class FakeEngine:
    """What SamplerLoop needs from evcdiff.engine.NCSNppEngine, without kernels."""
    split = False

    def __init__(self, B, H):
        self.device = torch.device("cpu")
        self.eps = torch.zeros(B, 15, H, H)
        self.xin = torch.zeros(B, H, H, 64, dtype=torch.bfloat16)
        self.ss_table = torch.zeros(1)
        self.labels, self.cond, self.cur, self.seen = None, None, None, []

    def set_labels(self, uniq):
        self.labels = list(uniq)

    def load_input(self, x, cond):
        self.cur, self.cond = x, cond

    def forward(self, idx, eps_out=None):
        self.seen.append(float(self.labels[idx]))
        (self.eps if eps_out is None else eps_out).copy_(_fake_eps(self.cur, self.labels[idx], self.cond))

    def refresh_x(self, x):
        pass


def _install(monkeypatch, eng):
    from evcdiff import models as M
    from evcdiff.models import loop as L

    def sampler_update(x, eps, noise, x_out, xin, k):  # csrc/sampler.cu: step_value
        if k.mode == 0:
            x0 = k.k0 * (x - k.k1 * eps)
            if k.clip:
                x0 = x0.clamp(-1, 1)
            r = k.c_x0 * x0 + k.c_x * x
            if k.c_eps != 0.0:
                r = r + k.c_eps * eps
            if k.c_noise != 0.0:
                r = r + k.c_noise * noise
        else:
            r = x - k.k1 * eps
        x_out.copy_(r)
        if xin is not None:
            eng.cur = x_out

    def pndm_update(x, es, x_out, et_out, xin, k):  # csrc/sampler.cu: pndm_value
        et = k.w[0] * es[0]
        for j in range(1, k.n_e):
            et = et + k.w[j] * es[j]
        et = et * k.w_scale
        r = x + k.d * (k.p * x - k.q * et)
        if k.clip:
            r = r.clamp(-1, 1)
        x_out.copy_(r)
        if xin is not None:
            eng.cur = x_out

    monkeypatch.setattr(L.ops, "sampler_update", sampler_update)
    monkeypatch.setattr(L.ops, "pndm_update", pndm_update)
    monkeypatch.setattr(L.torch.cuda, "device", lambda dev: contextlib.nullcontext())
    net = type("Net", (), {})()
    holder = {}

    def get(net_, B, device, precision=None):
        if "loop" not in holder:
            holder["loop"] = L.SamplerLoop(net_, eng)
        return holder["loop"]
    monkeypatch.setattr(M.SamplerLoop, "get", staticmethod(get))
    return M


def _net(cfg):
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    return UNetMore_DDPM(cfg)  # parameters unused: only the schedule buffers and the `engine` attribute matter


@pytest.mark.parametrize("kind,subsample,denoise,clip", [("ddpm", 10, True, True), ("ddpm", 100, True, True),
                                                        ("ddpm", 7, False, False), ("ddim", 10, True, True),
                                                        ("ddim", 25, False, True), ("ddim", 1000, True, True)])
def test_ancestral_plan_equals_oracle(monkeypatch, kind, subsample, denoise, clip):
    cfg = common.tiny_config()
    B, H = 2, 8
    eng = FakeEngine(B, H)
    M = _install(monkeypatch, eng)
    net = _net(cfg)
    g = torch.Generator().manual_seed(3)
    x_T = torch.randn(B, 15, H, H, generator=g)
    cond = torch.rand(B, 6, H, H, generator=g, dtype=torch.float64) * 2 - 1
    model = lambda x, labels: _fake_eps(x, float(labels[0]), cond)
    sched = S.schedule(cfg)
    if kind == "ddpm":
        n = len(S._subsample(sched[1], sched[2], sched[0], subsample)[0])
        tape = [torch.randn(B, 15, H, H, generator=g) for _ in range(n - 1)]
        ref = S.ddpm_sampler(x_T.clone(), model, sched, lambda i: tape[i], subsample_steps=subsample, denoise=denoise,
                             clip_before=clip)
        got = M.ddpm_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=denoise, subsample_steps=subsample,
                             clip_before=clip, noise=tape, graph=False)
    else:
        ref = S.ddim_sampler(x_T.clone(), model, sched, subsample_steps=subsample, denoise=denoise, clip_before=clip)
        got = M.ddim_sampler(x_T.clone(), net, cond=cond, final_only=True, denoise=denoise, subsample_steps=subsample,
                             clip_before=clip, graph=False)
    assert got.shape == (1, B, 15, H, H) and got.dtype == torch.float32
    assert common.rel_l2(got[0], ref) < 2e-6, common.rel_l2(got[0], ref)
    # labels the network saw: every schedule index, then the reference's denoise quirk (label L-1, not the last step)
    L = len(range(0, 1000, 1000 // subsample)) if subsample < 1000 else 1000
    want = [float(s) for s in (range(0, 1000, 1000 // subsample) if subsample < 1000 else range(1000))]
    assert eng.seen == want + ([float(L - 1)] if denoise else [])


@pytest.mark.parametrize("subsample,clip", [(20, True), (10, False), (5, True)])
def test_fpndm_plan_equals_oracle(monkeypatch, subsample, clip):
    cfg = common.tiny_config()
    B, H = 2, 8
    eng = FakeEngine(B, H)
    M = _install(monkeypatch, eng)
    net = _net(cfg)
    g = torch.Generator().manual_seed(4)
    x_T = torch.randn(B, 15, H, H, generator=g)
    cond = torch.rand(B, 6, H, H, generator=g, dtype=torch.float64) * 2 - 1
    seen = []
    model = lambda x, labels: _fake_eps(x, float(labels[0]), cond)
    ref = S.fpndm_sampler(x_T.clone(), model, S.schedule(cfg), subsample, clip_before=clip, labels_seen=seen)
    got = M.FPNDM_sampler(x_T.clone(), net, cond=cond, final_only=True, subsample_steps=subsample, clip_before=clip,
                          graph=False)
    assert common.rel_l2(got[0], ref) < 5e-6, common.rel_l2(got[0], ref)
    assert eng.seen == seen and len(seen) == 3 * 4 + (subsample - 3)  # 3 Runge-Kutta steps of 4 evaluations, then one each
    # per-step images when final_only=False: one entry per sampler step, on the host, like the reference's stack of x.to('cpu')
    eng2 = FakeEngine(B, H)
    M = _install(monkeypatch, eng2)
    allx = M.FPNDM_sampler(x_T.clone(), net, cond=cond, final_only=False, subsample_steps=subsample, clip_before=clip)
    assert allx.shape == (subsample, B, 15, H, H) and torch.equal(allx[-1], got[0])


def test_generate_frame_micro_batches_host_logic(monkeypatch):
    """pipeline.generate_frame (city_sender.py:326-351) on CPU with a recording sampler: data_transform 2x-1 on the
    conditioning frames (dtype kept, like the reference's float64 input), per-micro-batch slices of the caller's x_T and of
    every noise draw, inverse transform + clamp, (B, 5, 3, H, W) result -- whatever max_batch is."""
    from evcdiff import pipeline as P
    cfg = common.tiny_config()
    H = cfg.data.image_size
    monkeypatch.setattr(P.ops, "inverse_transform", lambda x, out: out.copy_(((x + 1.0) / 2.0).clamp(0, 1)))
    calls = []

    def sampler(x_T, net, cond=None, noise=None, **kw):
        calls.append((x_T.shape[0], cond.dtype, kw.get("subsample_steps"), kw.get("final_only")))
        assert cond.shape[0] == x_T.shape[0] and all(n.shape[0] == x_T.shape[0] for n in noise)
        out = 0.5 * x_T + 0.25 * cond.float().mean(dim=1, keepdim=True) + sum(noise)
        return out.unsqueeze(0)

    net = torch.nn.Linear(1, 1)  # only `next(net.parameters()).device` is used by the glue
    g = torch.Generator().manual_seed(5)
    B = 7
    frames_in = torch.rand(B, 6, H, H, generator=g, dtype=torch.float64)
    x_T = torch.randn(B, 15, H, H, generator=g)
    noise = [0.01 * torch.randn(B, 15, H, H, generator=g) for _ in range(3)]
    one = P.generate_frame(net, frames_in, config=cfg, sampler=sampler, init_samples=x_T, noise=noise, max_batch=64)
    assert one.shape == (B, cfg.data.num_frames, cfg.data.channels, H, H) and one.device.type == "cpu"
    assert float(one.min()) >= 0.0 and float(one.max()) <= 1.0
    want = ((0.5 * x_T + 0.25 * (2 * frames_in - 1).float().mean(dim=1, keepdim=True) + sum(noise) + 1) / 2).clamp(0, 1)
    assert torch.allclose(one.reshape(B, 15, H, H), want, atol=1e-6)
    assert calls == [(B, torch.float64, cfg.sampling.subsample, True)]
    calls.clear()
    parts = P.generate_frame(net, frames_in, config=cfg, sampler=sampler, init_samples=x_T, noise=noise, max_batch=3)
    assert [c[0] for c in calls] == [3, 3, 1] and torch.equal(parts, one)
