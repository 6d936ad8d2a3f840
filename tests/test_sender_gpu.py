"""GPU tests of the batched sender loop: float64 PSNR, accept-prefix decision, autoregressive cycles."""
import numpy as np
import pytest
import torch

import common
from oracle import ncsnpp as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cal_psnr(img1, img2, maxvalue=1.0):  # restatement of city_sender.py:255-258
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    mse = np.mean((img1 - img2) ** 2)
    return 10 * np.log10((maxvalue ** 2) / mse)


def test_psnr_and_accept_prefix():
    from evcdiff import ops
    g = torch.Generator(device=DEV).manual_seed(1)
    a = torch.rand(4, 5, 3, 32, 32, device=DEV, generator=g)
    b = (a + 0.05 * torch.randn(4, 5, 3, 32, 32, device=DEV, generator=g)).clamp(0, 1)
    p = ops.frame_psnr(a, b)
    ref = np.array([[cal_psnr(a[v, f].cpu().numpy(), b[v, f].cpu().numpy()) for f in range(5)] for v in range(4)])
    assert p.dtype == torch.float64 and np.allclose(p.cpu().numpy(), ref, rtol=1e-12, atol=1e-10)
    score = torch.tensor([[30, 31, 10, 40, 40], [5, 40, 40, 40, 40], [30, 30, 30, 30, 30], [30, 30, 30, 30, 29.9]],
                         dtype=torch.float64, device=DEV)
    assert ops.accept_prefix(score, 30.0).tolist() == [2, 0, 5, 4]
    assert ops.accept_prefix(-score, -30.0, higher_is_better=False).tolist() == [2, 0, 5, 4]


def test_batched_sender_cycles():
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    from evcdiff.sender import BatchedSender
    cfg = common.gpu64_config(device=DEV)
    cfg.sampling.subsample = 5
    net = UNetMore_DDPM(cfg)
    net.load_state_dict(common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=4, active=True), strict=False)
    net = net.to(DEV).eval()
    g = torch.Generator(device=DEV).manual_seed(2)
    V, T = 3, 12
    x_gt = torch.rand(V, T, 3, 32, 32, device=DEV, generator=g)
    # threshold -inf: every predicted frame is accepted -> 2 keyframes then 5 frames per cycle
    x_ge, d, n = BatchedSender(net, cfg, threshold=-1e30).encode(x_gt)
    assert n == 2 and x_ge.shape == x_gt.shape
    assert d.tolist() == [[1, 1] + [0] * 10] * V
    assert torch.equal(x_ge[:, :2], x_gt[:, :2]) and float(x_ge.min()) >= 0.0 and float(x_ge.max()) <= 1.0
    # threshold +inf: nothing is ever accepted -> every frame is a keyframe, two per cycle
    x_ge, d, n = BatchedSender(net, cfg, threshold=1e30).encode(x_gt)
    assert n == 5 and torch.equal(x_ge, x_gt) and int(d.min()) == 1
    # mixed: accept only frames with PSNR above the median PSNR of a probe cycle; invariants of the flag array
    sender = BatchedSender(net, cfg, threshold=7.0)
    x_ge, d, n = sender.encode(x_gt)
    assert x_ge.shape == x_gt.shape and set(d.unique().tolist()) <= {0, 1}
    key = d.bool()
    assert torch.equal(x_ge[key], x_gt[key])  # keyframes are the (lossless stand-in) ground truth


def _fake_predictor(cond01):
    """Deterministic stand-in for the diffusion sampler: predicted frame j = the last conditioning frame pulled towards
    mid-grey by a factor that grows with j and with the frame's first pixel, so that the PSNR against a slowly drifting
    ground truth falls with j at a video-dependent rate.  A function of the conditioning frames alone: the reference
    restatement and the batched sender see identical predictions as long as their reconstructions agree."""
    B, _, H, W = cond01.shape
    last = cond01[:, 3:6].float()
    m = last[:, :1, :1, :1]  # a per-video scalar without a reduction: bit-identical for any batch size
    frames = [(last + (0.5 - last) * (0.015 + 0.08 * m) * (j + 1)).clamp(0, 1) for j in range(5)]
    return torch.stack(frames, 1)  # (B, 5, 3, H, W)


def _fake_distance(p, g):
    """Stand-in for loss_fn_alex (city_sender.py:302,389; the AlexNet LPIPS weights are not in the image): a distance that
    is 0 for equal frames and grows with the difference, float64 so that the batched and the per-frame evaluation agree."""
    return (p.double() - g.double()).abs().mean(dim=(-3, -2, -1))


@pytest.mark.parametrize("rule", ["psnr", "lpips"])
@pytest.mark.parametrize("compact", [False, True])
def test_batched_sender_equals_reference_loop_per_video(monkeypatch, compact, rule):
    """BatchedSender against oracle/sender.py (restatement of SenderCity.update / decide_5to5 / decide_5to5_lpips and the
    driver's while loop, city_sender.py:353-437, 519-550), video by video: same flags d, same reconstruction x_ge, same
    number of cycles for the slowest video -- with and without compaction of finished videos, for the PSNR rule (score >=
    threshold) and the LPIPS rule (distance <= threshold, injected distance)."""
    from evcdiff import sender as S
    from oracle import sender as RS
    calls = []

    def fake_generate_frame(net, cond, **kw):
        calls.append(cond.shape[0])
        return _fake_predictor(cond)
    monkeypatch.setattr(S, "generate_frame", fake_generate_frame)
    g = torch.Generator(device=DEV).manual_seed(7)
    V, T, H = 11, 30, 32
    base = torch.rand(V, 1, 3, H, H, device=DEV, generator=g)
    drift = torch.linspace(0.002, 0.03, V, device=DEV).view(V, 1, 1, 1, 1) * torch.arange(T, device=DEV).view(1, T, 1, 1, 1)
    x_gt = (base * (0.4 + 0.6 * torch.rand(V, 1, 1, 1, 1, device=DEV, generator=g)) + drift).clamp(0, 1)
    cfg = common.gpu64_config(device=DEV)
    if rule == "psnr":
        thr, kw, lp = 31.0, {}, None
    else:
        thr, kw, lp = 0.02, dict(score_fn=_fake_distance, higher_is_better=False), (lambda p, g: float(_fake_distance(p, g)))
    snd = S.BatchedSender(None, cfg, threshold=thr, compact=compact, bucket=4, **kw)
    x_ge, d, n = snd.encode(x_gt)
    x_ge, d = x_ge.cpu(), d.cpu().numpy()
    cycles = []
    for v in range(V):
        gf = lambda frames: _fake_predictor(frames.to(DEV)).cpu()
        r_ge, r_d, r_n = RS.encode_video(x_gt[v].cpu(), gf, thr, total=T, lpips_fn=lp)
        assert d[v].tolist() == r_d.tolist(), (v, d[v].tolist(), r_d.tolist())
        assert torch.equal(x_ge[v], r_ge.float()), v
        cycles.append(r_n)
    assert n == max(cycles)
    assert len(set(cycles)) > 1 and 0 < int(d.sum()) < V * T  # the case really mixes keyframes and predictions
    if compact:
        assert snd.sampled_videos < V * n and all(c % 4 == 0 or c == V for c in calls)
    else:
        assert snd.sampled_videos == V * n
