"""CPU test of the HOST logic of the batched sender loop (evcdiff/sender.py: cycle bookkeeping, accepted-prefix scatter,
keyframe spending, compaction of finished videos) against oracle/sender.py, the restatement of city_sender.py:353-437,
519-550.  The two device kernels the loop calls (float64 PSNR, accept-prefix) are replaced by torch stand-ins here -- they
have their own GPU parity tests in tests/test_sender_gpu.py; nothing below needs a GPU."""
import numpy as np
import pytest
import torch

import common
from test_sender_gpu import _fake_distance, _fake_predictor


def _psnr_torch(pred, gt):
    mse = ((pred.double() - gt.double()) ** 2).mean(dim=(-3, -2, -1))
    return 10.0 * torch.log10(1.0 / mse)


def _accept_prefix_torch(score, threshold, higher_is_better=True):
    ok = (score >= threshold) if higher_is_better else (score <= threshold)
    return ok.long().cumprod(dim=1).sum(dim=1).to(torch.int32)


@pytest.mark.parametrize("rule", ["psnr", "lpips"])
@pytest.mark.parametrize("compact", [False, True])
def test_batched_sender_host_logic_equals_reference_loop(monkeypatch, compact, rule):
    from evcdiff import sender as S
    from oracle import sender as RS
    calls, keyframe_shapes = [], []

    def fake_generate_frame(net, cond, **kw):
        calls.append(cond.shape[0])
        return _fake_predictor(cond)

    def keyframes(frames):  # lossless stand-in that records what the codec callback is handed
        keyframe_shapes.append(tuple(frames.shape))
        return frames
    monkeypatch.setattr(S, "generate_frame", fake_generate_frame)
    monkeypatch.setattr(S.ops, "frame_psnr", _psnr_torch)
    monkeypatch.setattr(S.ops, "accept_prefix", _accept_prefix_torch)
    g = torch.Generator().manual_seed(7)
    V, T, H = 11, 30, 16
    base = torch.rand(V, 1, 3, H, H, generator=g)
    drift = torch.linspace(0.002, 0.03, V).view(V, 1, 1, 1, 1) * torch.arange(T).view(1, T, 1, 1, 1)
    x_gt = (base * (0.4 + 0.6 * torch.rand(V, 1, 1, 1, 1, generator=g)) + drift).clamp(0, 1)
    cfg = common.gpu64_config(device="cpu")
    if rule == "psnr":
        thr, kw, lp = 31.0, {}, None
    else:
        thr, kw, lp = 0.02, dict(score_fn=_fake_distance, higher_is_better=False), (lambda p, q: float(_fake_distance(p, q)))
    snd = S.BatchedSender(None, cfg, threshold=thr, compact=compact, bucket=4, keyframe_fn=keyframes, **kw)
    x_ge, d, n = snd.encode(x_gt)
    d = d.numpy()
    cycles = []
    for v in range(V):
        r_ge, r_d, r_n = RS.encode_video(x_gt[v], _fake_predictor, thr, total=T, lpips_fn=lp)
        assert d[v].tolist() == r_d.tolist(), (v, d[v].tolist(), r_d.tolist())
        assert torch.equal(x_ge[v], r_ge.float()), v
        cycles.append(r_n)
    assert n == max(cycles)
    assert len(set(cycles)) > 1 and 0 < int(d.sum()) < V * T  # the case mixes keyframes and predictions
    assert all(len(s) == 4 and s[1:] == (3, H, H) for s in keyframe_shapes)  # the codec callback always sees (n, 3, H, W)
    if compact:
        assert snd.sampled_videos < V * n and all(c % 4 == 0 or c == V for c in calls)
    else:
        assert snd.sampled_videos == V * n
    # flags: 1 = coded keyframe, 0 = predicted; keyframes are the (lossless) ground truth
    key = torch.from_numpy(d.astype(np.bool_))
    assert torch.equal(x_ge[key], x_gt[:, :T][key])
