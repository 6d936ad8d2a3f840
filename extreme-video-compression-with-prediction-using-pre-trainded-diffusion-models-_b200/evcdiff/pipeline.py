"""The call a user of the reference makes for the hot path: SenderCity.generate_frame (city_sender.py:326-351)
without the per-call checkpoint reload -- conditioning frames in [0,1] in, predicted frames in [0,1] out."""
import torch

from . import ops
from .models import FPNDM_sampler, ddim_sampler, ddpm_sampler

SAMPLERS = {"DDPM": ddpm_sampler, "DDIM": ddim_sampler, "FPNDM": FPNDM_sampler}


def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of `n_items` videos for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def get_model(config, ckpt_file=None, states=None, kind="ncsnpp"):
    """SenderCity.get_model (city_sender.py:304-324), done ONCE instead of once per sampling cycle: build the model,
    load the checkpoint (`states[0]` = state dict whose keys may carry DataParallel's 'module.' prefix, `states[-1]` =
    EMA shadow, used when config.model.ema), return the evaluation-mode model on config.device.  The bf16 operand
    repack happens lazily at the first sampling call and again only if the parameters change."""
    from .models.ema import EMAHelper
    if kind == "ncsnpp":
        from .models.better.ncsnpp_more import UNetMore_DDPM as Model
    else:
        from .models.unet import UNet_DDPM as Model
    net = Model(config).to(config.device)
    if states is None and ckpt_file is not None:
        states = torch.load(ckpt_file, map_location=config.device)
    if states is not None:
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in states[0].items()}
        net.load_state_dict(sd, strict=False)
        if getattr(config.model, "ema", False):
            helper = EMAHelper(mu=config.model.ema_rate)
            helper.register(net)
            helper.load_state_dict({(k[len("module."):] if k.startswith("module.") else k): v
                                    for k, v in states[-1].items()})
            helper.ema(net)
    return net.eval()


@torch.no_grad()
def generate_frame(net, input_frames, config=None, sampler="DDPM", init_samples=None, to_host=True,
                   max_batch=64, noise=None, **sampler_kwargs):
    """input_frames: (B, num_frames_cond*3, H, W) in [0,1], host (ideally pinned) or device, fp32/fp64.
    Returns predicted frames (B, num_frames, 1?, ...) as the reference does: (B, 5, 3, H, W) in [0,1] --
    on the host when to_host (the reference's pred.to('cpu')), else on the device.

    data_transform (2x-1), the sampler and inverse_data_transform (+clamp) all run on the GPU; batches larger
    than `max_batch` are processed in micro-batches that reuse one captured graph.  `init_samples` (B,15,H,W) and
    `noise` (n_draws,B,15,H,W) replace the generator draws of city_sender.py:330-333 / models/__init__.py:326 (used
    by the multi-GPU path to hand every rank its slice of the global-batch draws, and by the parity tests)."""
    config = config or net.config
    dev = next(net.parameters()).device
    d = config.data
    H = d.image_size
    c_x = d.channels * d.num_frames
    sk = dict(final_only=True, denoise=getattr(config.sampling, "denoise", True),
              subsample_steps=getattr(config.sampling, "subsample", None),
              clip_before=getattr(config.sampling, "clip_before", True), verbose=True, log=True)
    sk.update(sampler_kwargs)
    fn = SAMPLERS[sampler.upper()] if isinstance(sampler, str) else sampler
    B = input_frames.shape[0]
    outs = []
    for lo in range(0, B, max_batch):
        hi = min(B, lo + max_batch)
        cond = input_frames[lo:hi].to(dev, non_blocking=True)
        cond = 2 * cond - 1.0  # data_transform (function.py:62-63); stays fp64 if the input was fp64, like the reference
        if init_samples is None:
            x_T = torch.randn((hi - lo, c_x, H, H), device=dev)  # city_sender.py:330-333
        else:
            x_T = init_samples[lo:hi].to(dev)
        if noise is not None:
            sk["noise"] = [noise[i][lo:hi] for i in range(len(noise))]
        x0 = fn(x_T, net, cond=cond, **sk)[-1]
        frames = torch.empty_like(x0)
        ops.inverse_transform(x0.contiguous(), frames)
        outs.append(frames)
    frames = torch.cat(outs, 0) if len(outs) > 1 else outs[0]
    frames = frames.reshape(B, d.num_frames, d.channels, H, H)
    return frames.to("cpu") if to_host else frames


def global_draws(n_videos, config, sampler, subsample, seed, device):
    """The Gaussian draws the reference makes for a batch of `n_videos` on one GPU after seeding torch with `seed`
    (city_sender.py:52,216-219): init_samples = randn(B,15,H,W) (city_sender.py:330-333), then one randn_like per
    non-final DDPM step (models/__init__.py:326).  Returns (x_T, noise or None); DDIM / F-PNDM draw no step noise."""
    d = config.data
    shape = (n_videos, d.channels * d.num_frames, d.image_size, d.image_size)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    x_T = torch.randn(shape, device=device, generator=g)
    noise = None
    if str(sampler).upper() == "DDPM":
        T = config.model.num_classes
        steps = T if (subsample is None or subsample >= T) else len(range(0, T, T // subsample))
        noise = [torch.randn(shape, device=device, generator=g) for _ in range(steps - 1)]
    return x_T, noise


@torch.no_grad()
def generate_frame_sharded(net, input_frames, rank, world, config=None, sampler="DDPM", seed=1234, max_batch=64,
                           gather="fp32", group=None, draws_fn=global_draws, frame_fn=generate_frame, **sampler_kwargs):
    """Multi-GPU form of generate_frame: `input_frames` holds the GLOBAL batch (V, 6, H, W) in [0,1];
    rank r samples the contiguous shard shard_range(V, r, world) with ITS SLICE of the draws a single GPU would make
    for all V videos from `seed`, so the gathered result equals the unsharded one.  No collective inside the sampling
    loop; one gather of the predicted frames to rank 0 at the end (fp32, or uint8 = round(255 x) for 4x less NVLink
    traffic).  Returns (V, 5, 3, H, W) on rank 0 (device tensor; uint8 if gather == 'uint8'), None elsewhere.
    `draws_fn` / `frame_fn` are injection points for the CPU (gloo) test of this host logic."""
    import torch.distributed as dist
    config = config or net.config
    V = input_frames.shape[0]
    lo, hi = shard_range(V, rank, world)
    dev = next(net.parameters()).device if hasattr(net, "parameters") else input_frames.device
    subsample = sampler_kwargs.get("subsample_steps", getattr(config.sampling, "subsample", None))
    x_T, noise = draws_fn(V, config, sampler, subsample, seed, dev)
    local = None
    if hi > lo:
        local = frame_fn(net, input_frames[lo:hi], config=config, sampler=sampler, init_samples=x_T[lo:hi],
                         noise=None if noise is None else [n[lo:hi] for n in noise], to_host=False, max_batch=max_batch,
                         **sampler_kwargs)
    del x_T, noise
    if gather == "uint8" and local is not None:
        local = frames_to_uint8(local)
    if world == 1:
        return local
    # equal-size buffers for dist.gather: shards differ by at most one video
    sizes = [shard_range(V, r, world) for r in range(world)]
    pad = max(h - l for l, h in sizes)
    d = config.data
    shape = (pad, d.num_frames, d.channels, d.image_size, d.image_size)
    dtype = torch.uint8 if gather == "uint8" else torch.float32
    buf = torch.zeros(shape, dtype=dtype, device=dev)
    if local is not None:
        buf[: hi - lo] = local
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0, group=group)
    if rank != 0:
        return None
    return torch.cat([out[r][: sizes[r][1] - sizes[r][0]] for r in range(world)])


def frames_to_uint8(frames):
    """[0,1] fp32 frames -> uint8 round(255 x) on the device (the format city_bonn.npy stores, city_sender.py:487)."""
    if frames.is_cuda:
        out = torch.empty(frames.shape, dtype=torch.uint8, device=frames.device)
        ops.frames_to_uint8(frames.contiguous(), out)
        return out
    return (frames * 255.0).round().clamp(0, 255).to(torch.uint8)
