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
                   max_batch=64, **sampler_kwargs):
    """input_frames: (B, num_frames_cond*3, H, W) in [0,1], host (ideally pinned) or device, fp32/fp64.
    Returns predicted frames (B, num_frames, 1?, ...) as the reference does: (B, 5, 3, H, W) in [0,1] --
    on the host when to_host (the reference's pred.to('cpu')), else on the device.

    data_transform (2x-1), the sampler and inverse_data_transform (+clamp) all run on the GPU; batches larger
    than `max_batch` are processed in micro-batches that reuse one captured graph."""
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
        x0 = fn(x_T, net, cond=cond, **sk)[-1]
        frames = torch.empty_like(x0)
        ops.inverse_transform(x0.contiguous(), frames)
        outs.append(frames)
    frames = torch.cat(outs, 0) if len(outs) > 1 else outs[0]
    frames = frames.reshape(B, d.num_frames, d.channels, H, H)
    return frames.to("cpu") if to_host else frames
