"""The multi-step sampling loop on one GPU: static state buffers, kernel enqueue order, CUDA-graph capture.

One SamplerLoop per (model, batch size, device).  It owns
  x      (B,15,H,W) fp32   sampler state, the layout the reference samplers return
  noise  (B,15,H,W) fp32   filled by `normal_()` from the global CUDA generator (same draw as torch.randn_like)
  e[j]   (B,15,H,W) fp32   eps history / Runge-Kutta stages for F-PNDM
and the UNet engine (evcdiff.engine) whose `xin` (NHWC bf16) every update kernel refreshes in place.
A whole loop (e.g. 101 UNet evaluations + 101 updates + 99 noise draws for DDPM-100) is captured once into a
single CUDA graph per key and replayed; inputs are copied into the static buffers before the replay.
"""
import torch

from .. import _lib, ops
from .._lib import EvcError, PndmCoef


class SamplerLoop:
    @staticmethod
    def get(net, B, device, precision=None):
        device = torch.device(device)
        if device.type != "cuda":
            raise EvcError("evcdiff samplers need CUDA tensors (no CPU fallback)")
        eng = net.engine(B, device, precision=precision)
        loop = getattr(eng, "_loop", None)
        if loop is None:
            loop = SamplerLoop(net, eng)
            eng._loop = loop
        return loop

    def __init__(self, net, eng):
        self.net, self.eng = net, eng
        dev = eng.device
        shape = tuple(eng.eps.shape)
        self.x = torch.zeros(shape, dtype=torch.float32, device=dev)
        self.noise = torch.zeros(shape, dtype=torch.float32, device=dev)
        self.e = None
        self.scratch = None
        self.tape = None  # static per-step noise tape (caller-supplied noise inside a captured graph)
        self.graphs = {}
        self.launches_per_run = {}

    # ------------------------------------------------------------------------------------------
    def _load(self, x_mod, cond):
        if tuple(x_mod.shape) != tuple(self.x.shape):
            raise EvcError(f"x_mod shape {tuple(x_mod.shape)} does not match the model ({tuple(self.x.shape)})")
        self.x.copy_(x_mod.to(torch.float32))
        self.eng.xin.zero_()
        self.eng.load_input(self.x, cond)

    def _capture_or_run(self, key, body, graph):
        """body() enqueues the loop on the current stream.  With graph=True it is captured once and replayed."""
        # small batches are launch-latency bound: overlap kernel prologues with programmatic dependent launch (same-box A/B,
        # profiles/r02_notes.md: +4 % at 1 video, +0.4-1.1 % at 12, neutral at 23, -0.7 % at 46)
        pixels = self.x.shape[0] * self.x.shape[2] * self.x.shape[3]
        _lib.load().evc_set_pdl(1 if pixels <= 16 * 128 * 128 else 0)
        if not graph:
            n0 = ops.launch_count()
            body()
            self.launches_per_run[key] = ops.launch_count() - n0
            return
        g = self.graphs.get(key)
        if g is None:
            # warm-up run on a side stream (module loading, cudaFuncSetAttribute, lazy allocations) --
            # capture must not be the first execution.  The state is restored afterwards.
            x0, xin0 = self.x.clone(), self.eng.xin.clone()
            rng = torch.cuda.get_rng_state(self.eng.device)
            s = torch.cuda.Stream(device=self.eng.device)
            s.wait_stream(torch.cuda.current_stream(self.eng.device))
            with torch.cuda.stream(s):
                self.eng.forward(0)
                self._warm_updates()
            torch.cuda.current_stream(self.eng.device).wait_stream(s)
            torch.cuda.synchronize(self.eng.device)
            self.x.copy_(x0)
            self.eng.xin.copy_(xin0)
            torch.cuda.set_rng_state(rng, self.eng.device)
            g = torch.cuda.CUDAGraph()
            n0 = ops.launch_count()
            with torch.cuda.graph(g):
                body()
            self.launches_per_run[key] = ops.launch_count() - n0
            self.graphs[key] = g
            # capture does not execute: restore state (nothing ran) and replay below
        g.replay()

    def _warm_updates(self):
        """First execution of the update kernels / RNG kernel outside capture (results discarded)."""
        from .._lib import StepCoef
        if self.scratch is None:
            self.scratch = torch.zeros_like(self.x)
        self.noise.normal_()
        ops.sampler_update(self.x, self.eng.eps, self.noise, self.scratch, None, StepCoef(0, 1, 1.0, 0.5, 0.5, 0.5, 0.0, 0.1))
        c = PndmCoef()
        c.n_e, c.clip, c.w_scale, c.d, c.p, c.q = 1, 1, 1.0, 0.1, 0.1, 0.1
        c.w[0] = 1.0
        ops.pndm_update(self.x, [self.eng.eps], self.scratch, None, None, c)
        self.scratch.copy_(self.eng.eps)

    # ------------------------------------------------------------------------------------------
    def _check_noise(self, t, what):
        """Caller-supplied noise must be exactly what the kernel reads: fp32, contiguous, the shape of x, on this device."""
        if not torch.is_tensor(t):
            raise EvcError(f"{what} must be a tensor")
        if tuple(t.shape) != tuple(self.x.shape):
            raise EvcError(f"{what} has shape {tuple(t.shape)}, expected {tuple(self.x.shape)}")
        return t.to(self.x.device, torch.float32).contiguous()

    def run_ancestral(self, key, x_mod, cond, labels, coefs, final_only, noise=None, noise_const=None, graph=True,
                      draws=None):
        """draws[i]: does step i consume one Gaussian draw (the reference draws `randn_like` on every non-final DDPM
        step, models/__init__.py:313-326, whatever the value of the coefficient)."""
        with torch.cuda.device(self.eng.device):
            return self._run_ancestral(key, x_mod, cond, labels, coefs, final_only, noise, noise_const, graph, draws)

    def _run_ancestral(self, key, x_mod, cond, labels, coefs, final_only, noise, noise_const, graph, draws):
        eng = self.eng
        if draws is None:
            draws = [c.mode == 0 and c.c_noise != 0.0 for c in coefs]
        n_draws = sum(1 for d in draws if d)
        uniq = list(dict.fromkeys(labels))
        idx = [uniq.index(v) for v in labels]
        eng.set_labels(uniq)
        self._load(x_mod, cond)
        images = []
        if noise_const is not None:
            noise_const = self._check_noise(noise_const, "noise_val")
        if not final_only:
            graph = False  # per-step host interaction: eager launches
        tape = None
        if noise is not None and noise_const is None:
            if len(noise) < n_draws:
                raise EvcError(f"noise has {len(noise)} entries, the schedule draws {n_draws}")
            if graph:
                # static tape: the captured graph reads slot i at the i-th draw; contents are refreshed before each replay
                if self.tape is None or self.tape.shape[0] < n_draws:
                    self.tape = torch.empty((n_draws,) + tuple(self.x.shape), dtype=torch.float32, device=self.x.device)
                for i in range(n_draws):
                    self.tape[i].copy_(self._check_noise(noise[i], f"noise[{i}]"))
                tape = self.tape
                key = key + ("tape", self.tape.data_ptr())
            else:
                tape = [self._check_noise(noise[i], f"noise[{i}]") for i in range(n_draws)]
        elif noise_const is not None:
            key = key + ("const", noise_const.data_ptr())
            graph = False

        def body():
            ni = 0
            for i, c in enumerate(coefs):
                eng.forward(idx[i])
                nz = None
                if draws[i]:
                    if noise_const is not None:
                        nz = noise_const
                    elif tape is not None:
                        nz = tape[ni]
                        ni += 1
                    else:
                        self.noise.normal_()
                        nz = self.noise
                ops.sampler_update(self.x, eng.eps, nz, self.x, eng.xin, c)
                eng.refresh_x(self.x)
                if not final_only:
                    images.append(self.x.to("cpu"))

        self._capture_or_run(key + (len(uniq), eng.ss_table.data_ptr()), body, graph)
        if final_only:
            return self.x.clone().unsqueeze(0)
        return torch.stack(images)

    # ------------------------------------------------------------------------------------------
    def run_fpndm(self, key, x_mod, cond, steps, steps_next, alphas_old, clip_before, final_only, graph=True):
        """F-PNDM (reference models/__init__.py:79-100 + models/pndm.py:3-52)."""
        with torch.cuda.device(self.eng.device):
            return self._run_fpndm(key, x_mod, cond, steps, steps_next, alphas_old, clip_before, final_only, graph)

    def _run_fpndm(self, key, x_mod, cond, steps, steps_next, alphas_old, clip_before, final_only, graph):
        eng = self.eng
        if self.e is None:
            self.e = [torch.zeros_like(self.x) for _ in range(7)]  # ring of 4 + 3 Runge-Kutta stages
            self.scratch = torch.zeros_like(self.x)
        # label sequence exactly as the network sees it: t.long() values and the float midpoints
        plan = []  # (kind, t, t_next)
        n_ets = 0
        for t, tn in zip(steps, steps_next):
            plan.append(("ab" if n_ets > 2 else "rk", float(t), float(tn)))
            n_ets += 1
        labs = []
        for kind, t, tn in plan:
            labs += [t] if kind == "ab" else [t, (t + tn) / 2, (t + tn) / 2, tn]
        uniq = list(dict.fromkeys(labs))
        eng.set_labels(uniq)
        li = {v: uniq.index(v) for v in uniq}
        self._load(x_mod, cond)
        images = []
        if not final_only:
            graph = False

        def coef(t, tn, n_e, w, w_scale):
            at = alphas_old[int(torch.tensor(t).long()) + 1]
            an = alphas_old[int(torch.tensor(tn).long()) + 1]
            d = an - at
            p = 1 / (at.sqrt() * (at.sqrt() + an.sqrt()))
            q = 1 / (at.sqrt() * (((1 - an) * at).sqrt() + ((1 - at) * an).sqrt()))
            c = PndmCoef()
            c.n_e, c.clip = n_e, int(bool(clip_before))
            for j in range(4):
                c.w[j] = float(w[j]) if j < n_e else 0.0
            c.w_scale = float(torch.tensor(w_scale, dtype=torch.float32))
            c.d, c.p, c.q = float(d), float(p), float(q)
            return c

        def evaluate(label_idx, dst):
            """eps of the current UNet input -> dst: the final conv's epilogue writes the history slot directly (bf16
            mode); split-precision mode keeps the engine's own eps buffer and copies."""
            if eng.split:
                eng.forward(label_idx)
                dst.copy_(eng.eps)
            else:
                eng.forward(label_idx, eps_out=dst)

        def body():
            ring = []  # eps history buffers, most recent last
            free = list(self.e)
            for kind, t, tn in plan:
                if len(ring) == 4:
                    free.append(ring.pop(0))
                e1 = free.pop()
                if kind == "ab":
                    evaluate(li[t], e1)
                    ring.append(e1)
                    es = [ring[-1], ring[-2], ring[-3], ring[-4]]
                    ops.pndm_update(self.x, es, self.x, None, eng.xin, coef(t, tn, 4, (55.0, -59.0, 37.0, -9.0), 1 / 24))
                    eng.refresh_x(self.x)
                else:
                    tm = (t + tn) / 2
                    e2, e3, e4 = free[-1], free[-2], free[-3]
                    evaluate(li[t], e1)
                    ring.append(e1)
                    ops.pndm_update(self.x, [e1], self.scratch, None, eng.xin, coef(t, tm, 1, (1.0,), 1.0))
                    eng.refresh_x(self.scratch)
                    evaluate(li[tm], e2)
                    ops.pndm_update(self.x, [e2], self.scratch, None, eng.xin, coef(t, tm, 1, (1.0,), 1.0))
                    eng.refresh_x(self.scratch)
                    evaluate(li[tm], e3)
                    ops.pndm_update(self.x, [e3], self.scratch, None, eng.xin, coef(t, tn, 1, (1.0,), 1.0))
                    eng.refresh_x(self.scratch)
                    evaluate(li[tn], e4)
                    ops.pndm_update(self.x, [e1, e2, e3, e4], self.x, None, eng.xin,
                                    coef(t, tn, 4, (1.0, 2.0, 2.0, 1.0), 1 / 6))
                    eng.refresh_x(self.x)
                if not final_only:
                    images.append(self.x.to("cpu"))

        self._capture_or_run(key + (len(uniq), eng.ss_table.data_ptr()), body, graph)
        if final_only:
            return self.x.clone().unsqueeze(0)
        return torch.stack(images)
