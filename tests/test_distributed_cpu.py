"""world_size-2 gloo test of the multi-GPU host logic: shard-by-video, per-rank noise slicing from the global
seed, and the single gather of predicted frames at the end.  No CUDA involved."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import common  # noqa: F401  (sets sys.path through conftest)


def _fake_sampler(cond, noise):
    # any per-video function: the path has no cross-sample op, so sharded == unsharded by construction
    return torch.tanh(cond.float().mean(dim=1, keepdim=True).repeat(1, 15, 1, 1) * 0.5 + 0.1 * noise)


def _worker(rank, world, port, n_videos, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evcdiff.pipeline import shard_range
    g = torch.Generator().manual_seed(1234)
    cond = torch.rand(n_videos, 6, 8, 8, generator=g, dtype=torch.float64)
    noise = torch.randn(n_videos, 15, 8, 8, generator=g)  # global-batch noise, sliced per rank
    lo, hi = shard_range(n_videos, rank, world)
    local = _fake_sampler(cond[lo:hi], noise[lo:hi])
    sizes = [shard_range(n_videos, r, world) for r in range(world)]
    pad = max(h - l for l, h in sizes)
    buf = torch.zeros(pad, *local.shape[1:])
    buf[: hi - lo] = local
    gathered = [torch.zeros_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, gathered, dst=0)
    if rank == 0:
        full = torch.cat([gathered[r][: sizes[r][1] - sizes[r][0]] for r in range(world)])
        torch.save({"full": full, "ref": _fake_sampler(cond, noise)}, out)
    dist.barrier()
    dist.destroy_process_group()


def _fake_frames(net, input_frames, config=None, sampler="DDPM", init_samples=None, noise=None, to_host=False,
                 max_batch=64, **kw):
    """Stand-in for pipeline.generate_frame with the same signature: any per-video function of (cond, x_T, noise)."""
    x = init_samples + 0.25 * input_frames.float().mean(dim=1, keepdim=True)
    for n in (noise or []):
        x = torch.tanh(x + 0.1 * n)
    return x.reshape(x.shape[0], config.data.num_frames, config.data.channels, *x.shape[-2:])


def _sharded_worker(rank, world, port, n_videos, sampler, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from evcdiff import pipeline
    cfg = common.make_config(image_size=8)
    cfg.sampling.subsample = 5
    g = torch.Generator().manual_seed(7)
    frames01 = torch.rand(n_videos, 6, 8, 8, generator=g, dtype=torch.float64)
    got = pipeline.generate_frame_sharded(torch.nn.Linear(1, 1), frames01, rank, world, config=cfg, sampler=sampler,
                                          seed=1234, frame_fn=_fake_frames, gather="fp32")
    got8 = pipeline.generate_frame_sharded(torch.nn.Linear(1, 1), frames01, rank, world, config=cfg, sampler=sampler,
                                           seed=1234, frame_fn=lambda *a, **k: torch.sigmoid(_fake_frames(*a, **k)),
                                           gather="uint8")
    if rank == 0:
        ref = pipeline.generate_frame_sharded(torch.nn.Linear(1, 1), frames01, 0, 1, config=cfg, sampler=sampler,
                                              seed=1234, frame_fn=_fake_frames)
        torch.save({"got": got, "ref": ref, "got8": got8}, out)
    else:
        assert got is None and got8 is None
    dist.barrier()
    dist.destroy_process_group()


def test_generate_frame_sharded_two_ranks_equals_one(tmp_path):
    """The real host logic of the multi-GPU path (pipeline.generate_frame_sharded: global-seed draws sliced per rank,
    uneven shards padded for the gather, fp32 and uint8 gathers) with a stand-in for the CUDA sampler."""
    for i, (n, sampler) in enumerate([(5, "DDPM"), (1, "DDPM"), (4, "FPNDM")]):
        out = str(tmp_path / f"res{i}.pt")
        port = 29600 + (os.getpid() % 2000) + i
        mp.spawn(_sharded_worker, args=(2, port, n, sampler, out), nprocs=2, join=True)
        res = torch.load(out)
        assert res["got"].shape == (n, 5, 3, 8, 8)
        assert torch.equal(res["got"], res["ref"])
        assert res["got8"].dtype == torch.uint8 and res["got8"].shape == (n, 5, 3, 8, 8)
        assert torch.equal(res["got8"], (torch.sigmoid(res["ref"]) * 255).round().to(torch.uint8))


def test_global_draws_match_reference_rng_order():
    """x_T first, then one draw per non-final DDPM step, from one generator: the order of city_sender.py:330-333 and
    models/__init__.py:326; DDIM / F-PNDM draw no step noise."""
    from evcdiff import pipeline
    cfg = common.make_config(image_size=8)
    x_T, noise = pipeline.global_draws(3, cfg, "DDPM", 10, 99, "cpu")
    g = torch.Generator().manual_seed(99)
    assert torch.equal(x_T, torch.randn(3, 15, 8, 8, generator=g))
    assert len(noise) == 9 and torch.equal(noise[0], torch.randn(3, 15, 8, 8, generator=g))
    assert pipeline.global_draws(3, cfg, "DDIM", 10, 99, "cpu")[1] is None
    assert len(pipeline.global_draws(2, cfg, "DDPM", None, 1, "cpu")[1]) == 999


def test_shard_ranges():
    from evcdiff.pipeline import shard_range
    assert [shard_range(46, r, 8) for r in range(8)] == [(0, 6), (6, 12), (12, 18), (18, 24), (24, 30), (30, 36),
                                                         (36, 41), (41, 46)]
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert shard_range(1, 1, 2) == (1, 1)


def test_two_rank_gather_equals_unsharded(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 5, out), nprocs=2, join=True)
    res = torch.load(out)
    assert torch.equal(res["full"], res["ref"])


def test_config_loader_matches_test_config():
    from evcdiff.config import load_config
    cfg = load_config(device="cpu", config_mod=["sampling.subsample=20 model.version=FPNDM"])
    ref = common.full_config()
    assert cfg.sampling.subsample == 20 and cfg.model.version == "FPNDM"
    for k in ("ngf", "ch_mult", "num_res_blocks", "attn_resolutions", "n_head_channels", "sigma_begin", "sigma_end",
              "num_classes", "sigma_dist", "spade", "cond_emb", "noise_in_cond", "gamma"):
        assert getattr(cfg.model, k) == getattr(ref.model, k), k
    for k in ("image_size", "channels", "num_frames", "num_frames_cond", "num_frames_future", "rescaled"):
        assert getattr(cfg.data, k) == getattr(ref.data, k), k
