"""Batched, GPU-resident version of the sender loop around the hot path:
SenderCity.update / decide_5to5 (city_sender.py:353-437) and the `while x_ge.shape[1] < 30` loop (:534-550).

The reference handles one video at a time and round-trips every cycle through numpy.  Here all videos of a shard
advance together: one captured sampling graph per cycle for the whole batch, PSNR and the accept-prefix decision
on the GPU, frames resident in HBM.  Keyframe coding (ELIC, out of scope) is a caller-supplied callback; the
default stand-in transmits the ground-truth frames unchanged.  The accept decision is decide_5to5 (PSNR >= threshold,
:353-374) by default; decide_5to5_lpips (:376-406: a perceptual distance <= threshold) is the same prefix rule with the
comparison reversed, selected with `score_fn=..., higher_is_better=False` -- the LPIPS network itself (lpips==0.1.4,
AlexNet weights, not in this image) is the caller's, e.g.
`score_fn=lambda pred, gt: loss_fn_alex(pred.flatten(0, 1), gt.flatten(0, 1)).view(pred.shape[:2])`.
"""
import torch

from . import ops
from .pipeline import generate_frame


class BatchedSender:
    def __init__(self, net, config=None, threshold=20.0, sampler="DDPM", keyframe_fn=None, max_batch=64, compact=True,
                 bucket=8, score_fn=None, higher_is_better=True, **sampler_kwargs):
        self.net = net
        self.config = config or net.config
        self.threshold = threshold
        # per-frame score (pred, gt: (V, 5, 3, H, W) in [0,1] on the device) -> (V, 5); None = float64 PSNR on the GPU
        # (cal_psnr, city_sender.py:255-258).  higher_is_better=False gives the `score <= threshold` rule of
        # decide_5to5_lpips (city_sender.py:392)
        self.score_fn = score_fn
        self.higher_is_better = bool(higher_is_better)
        self.sampler = sampler
        # (n, 3, H, W) frames in [0,1] -> decoded frames of the same shape; ELIC stand-in: lossless keyframes
        self.keyframe_fn = keyframe_fn or (lambda frames_gt: frames_gt)
        self.max_batch = max_batch
        # finished videos leave the sampling batch (the reference stops calling update() for a video that has its 30
        # frames, city_sender.py:534); the active set is padded to a multiple of `bucket` so that only a few batch
        # sizes (= engines / captured graphs) ever exist
        self.compact = compact
        self.bucket = max(1, int(bucket))
        self.sampled_videos = 0  # videos x cycles actually sampled (diagnostic)
        self.sampler_kwargs = sampler_kwargs
        self.num_frames = self.config.data.num_frames
        self.num_cond = self.config.data.num_frames_cond

    @torch.no_grad()
    def encode(self, x_gt, total_frames=None):
        """x_gt: (V, T, 3, H, W) in [0,1] on the device.  Returns (x_ge, d, n_cycles): reconstructed frames (V,T,3,H,W),
        the per-frame flag array d (1 = transmitted keyframe, 0 = predicted by the diffusion model; city_sender.py:530)
        and the number of sampling cycles executed."""
        V, T, C, H, W = x_gt.shape
        T = total_frames or T
        dev = x_gt.device
        x_gt = x_gt.float()
        x_ge = torch.zeros((V, T + self.num_frames + self.num_cond, C, H, W), dtype=torch.float32, device=dev)
        d = torch.ones((V, T + self.num_frames + self.num_cond), dtype=torch.int32, device=dev)
        pos = torch.full((V,), self.num_cond, dtype=torch.long, device=dev)  # frames available per video
        x_ge[:, :self.num_cond] = self.keyframe_fn(x_gt[:, :self.num_cond].reshape(-1, C, H, W)).reshape(
            V, self.num_cond, C, H, W)  # keyframe_fn always sees (n, 3, H, W) frames
        ar = torch.arange(V, device=dev)
        n_cycles = 0
        while bool((pos < T).any()):
            n_cycles += 1
            # conditioning = last num_cond reconstructed frames of every video
            idx = (pos.clamp(max=T)[:, None] - self.num_cond + torch.arange(self.num_cond, device=dev)[None, :])
            cond = x_ge[ar[:, None], idx].reshape(V, self.num_cond * C, H, W)
            active = (pos < T).nonzero().flatten()
            if self.compact and active.numel() < V:
                n_act = int(active.numel())
                n_run = min(V, -(-n_act // self.bucket) * self.bucket)
                run = torch.cat([active, active[-1:].expand(n_run - n_act)])  # padding repeats the last active video
                sub = generate_frame(self.net, cond[run], config=self.config, sampler=self.sampler, to_host=False,
                                     max_batch=self.max_batch, **self.sampler_kwargs)
                pred = torch.zeros((V,) + tuple(sub.shape[1:]), dtype=sub.dtype, device=dev)
                pred[active] = sub[:n_act]
                self.sampled_videos += n_run
            else:
                pred = generate_frame(self.net, cond, config=self.config, sampler=self.sampler, to_host=False,
                                      max_batch=self.max_batch, **self.sampler_kwargs)  # (V, 5, 3, H, W)
                self.sampled_videos += V
            gidx = (pos[:, None] + torch.arange(self.num_frames, device=dev)[None, :]).clamp(max=x_gt.shape[1] - 1)
            gt = x_gt[ar[:, None], gidx]
            if self.score_fn is None:
                score = ops.frame_psnr(pred.contiguous(), gt.contiguous())  # (V, 5) float64
            else:
                score = self.score_fn(pred, gt).to(torch.float64).reshape(V, self.num_frames)
            acc = ops.accept_prefix(score, self.threshold, higher_is_better=self.higher_is_better).long()
            acc = torch.where(pos < T, acc, torch.zeros_like(acc))
            for j in range(self.num_frames):  # accepted prefix -> reconstruction, flag 0
                take = (acc > j) & (pos + j < T)
                if bool(take.any()):
                    x_ge[ar[take], (pos + j)[take]] = pred[take, j]
                    d[ar[take], (pos + j)[take]] = 0
            none = (acc == 0) & (pos < T)  # nothing accepted: spend two more keyframes (city_sender.py:537-548)
            if bool(none.any()):
                for j in range(self.num_cond):
                    src = (pos + j).clamp(max=x_gt.shape[1] - 1)
                    x_ge[ar[none], (pos + j)[none]] = self.keyframe_fn(x_gt[ar[none], src[none]])
            pos = pos + torch.where(acc > 0, acc, torch.where(pos < T, torch.full_like(acc, self.num_cond), torch.zeros_like(acc)))
        return x_ge[:, :T], d[:, :T], n_cycles
