"""Two-GPU parity of the sharded sampling path (VERDICT r01 'missing' item 2): the frames gathered from
two NCCL ranks -- each sampling its shard with its slice of the global-seed draws -- equal the single-GPU result for
the whole batch: bit for bit when the single GPU uses the shard sizes as micro-batches (same tile configuration), and
up to bf16 rounding noise when it samples all videos in one batch.  Needs two CUDA devices (`gpurun --gpus 2 -- python -m pytest tests/test_two_gpu.py -m gpu`);
skipped on a one-GPU box."""
import os

import pytest
import torch
import torch.multiprocessing as mp

import common

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, sampler, subsample, n_videos, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from evcdiff import pipeline
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    from oracle import ncsnpp as O
    cfg = common.full_config(device=dev)
    cfg.sampling.subsample = subsample
    net = UNetMore_DDPM(cfg)
    net.load_state_dict(common.seeded_state_dict(O.ncsnpp_param_shapes(cfg), seed=9, active=True), strict=False)
    net = net.to(dev).eval()
    g = torch.Generator().manual_seed(3)
    frames01 = torch.rand(n_videos, 6, 128, 128, generator=g, dtype=torch.float64)
    got = pipeline.generate_frame_sharded(net, frames01, rank, world, config=cfg, sampler=sampler, seed=1234)
    got8 = pipeline.generate_frame_sharded(net, frames01, rank, world, config=cfg, sampler=sampler, seed=1234, gather="uint8")
    if rank == 0:
        ref = pipeline.generate_frame_sharded(net, frames01, 0, 1, config=cfg, sampler=sampler, seed=1234)
        # the same videos on one GPU in the shard-sized batches (3 + 2): identical tile configuration -> identical bits
        x_T, noise = pipeline.global_draws(n_videos, cfg, sampler, subsample, 1234, dev)
        parts = []
        for r in range(world):
            lo, hi = pipeline.shard_range(n_videos, r, world)
            parts.append(pipeline.generate_frame(net, frames01[lo:hi], config=cfg, sampler=sampler, init_samples=x_T[lo:hi],
                                                 noise=None if noise is None else [n[lo:hi] for n in noise], to_host=False))
        torch.save({"got": got.cpu(), "ref": ref.cpu(), "got8": got8.cpu(), "same_tiles": torch.cat(parts).cpu()}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("sampler,subsample", [("DDPM", 5), ("FPNDM", 5)])
def test_sharded_equals_unsharded_on_two_gpus(tmp_path, sampler, subsample):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    out = str(tmp_path / "res.pt")
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, sampler, subsample, 5, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["got"].shape == (5, 5, 3, 128, 128)
    assert torch.equal(res["got"], res["same_tiles"]), common.rel_l2(res["got"], res["same_tiles"])
    # one batch of 5 on one GPU: another tile configuration, i.e. other bf16 rounding noise.  DDPM washes it out with
    # fresh noise (measured < 1e-2); F-PNDM without noise is a chaotic map for an untrained network and amplifies it
    # (measured 1.6e-2 after 5 coarse steps)
    tol = 1e-2 if sampler == "DDPM" else 5e-2
    assert common.rel_l2(res["got"], res["ref"]) < tol, common.rel_l2(res["got"], res["ref"])
    assert torch.equal(res["got8"], (res["got"] * 255.0).round().to(torch.uint8))
