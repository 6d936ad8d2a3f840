"""CPU tests of the host side: the C-ABI library loads and exports every declared symbol, the model mirrors the
reference's state dict, unsupported configurations and CPU tensors fail loudly (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

import common
from oracle import ncsnpp as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from evcdiff import _lib
    hdr = open(os.path.join(ROOT, "include", "evcdiff.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(evc_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.evc_version() >= 1
    assert lib.evc_launch_count() == 0 or lib.evc_launch_count() > 0


def test_invalid_arguments_are_errors_not_fallbacks():
    from evcdiff import _lib
    lib = _lib.load()
    d = _lib.GemmDesc()
    h = ctypes.c_void_p()
    assert lib.evc_gemm_plan_create(ctypes.byref(d), ctypes.byref(h)) == -1
    assert b"n_seg" in lib.evc_last_error()
    assert lib.evc_gn_stats(None, 8, 1, 1, 8, None, 8, 0, None, 0, None) == -1
    assert lib.evc_softmax_rows(None, None, 1, 3, None) == -1


def test_state_dict_matches_reference_layout():
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM, ncsnpp_spec
    for cfgf in (common.tiny_config, common.gpu64_config):
        cfg = cfgf()
        net = UNetMore_DDPM(cfg)
        sd = net.state_dict()
        shapes = O.ncsnpp_param_shapes(cfg)
        keys = [k for k in sd if k not in ("betas", "alphas", "alphas_prev", "unet.sigmas")]
        assert keys == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in keys)
        assert ncsnpp_spec(cfg) == [{**s, **({"init_scale": 1.0} if i == 2 else {}),
                                     **({"init_scale": 0.0} if i == len(O.ncsnpp_spec(cfg)) - 1 else {})}
                                    for i, s in enumerate(O.ncsnpp_spec(cfg))]
        # schedule buffers: index 0 = noisiest
        assert torch.equal(net.alphas_prev[:-1], net.alphas[1:]) and float(net.alphas_prev[-1]) == 1.0
        assert float(net.alphas[0]) < 1e-4 and float(net.alphas[-1]) > 0.999
        # zero-init layers of the reference (Conv_1, NIN_3, output conv)
        assert float(sd["unet.all_modules.3.Conv_1.weight"].abs().max()) < 1e-4
        # DataParallel-style checkpoints load after stripping the `module.` prefix (city_sender.py:313-322)
        dp = {"module." + k: v for k, v in sd.items()}
        net2 = UNetMore_DDPM(cfg)
        net2.load_state_dict({k[len("module."):]: v for k, v in dp.items()}, strict=True)


def test_no_cpu_fallback():
    from evcdiff import models as M
    from evcdiff._lib import EvcError
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    cfg = common.tiny_config()
    net = UNetMore_DDPM(cfg)
    x = torch.randn(1, 15, 16, 16)
    with pytest.raises(EvcError):
        net(x, torch.zeros(1, dtype=torch.long), cond=torch.zeros(1, 6, 16, 16))
    with pytest.raises(EvcError):
        M.ddpm_sampler(x, net, cond=None, subsample_steps=10)
    with pytest.raises(EvcError):
        M.ddpm_sampler(x, torch.nn.Identity(), subsample_steps=10)
    with pytest.raises(EvcError):
        M.ddpm_sampler(x, net, gamma=True)
    with pytest.raises(EvcError):
        M.ddim_sampler(x, net, t_min=0.5)
    with pytest.raises(TypeError):
        M.FPNDM_sampler(x, net, subsample_steps=None)  # reference models/__init__.py:62 crashes the same way


def test_unsupported_configs_raise():
    from evcdiff._lib import EvcError
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    for key in ("spade", "cond_emb", "noise_in_cond", "gamma"):
        cfg = common.tiny_config()
        setattr(cfg.model, key, True)
        with pytest.raises(EvcError):
            UNetMore_DDPM(cfg)
    cfg = common.tiny_config()
    cfg.model.arch = "unetmore3d"
    with pytest.raises(EvcError):
        UNetMore_DDPM(cfg)


def test_get_sigmas_and_coefficients():
    from evcdiff import models as M
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    from oracle import samplers as S
    cfg = common.tiny_config()
    assert torch.equal(M.get_sigmas(cfg), S.schedule(cfg)[0])
    net = UNetMore_DDPM(cfg)
    steps, a, ap, b = M._subsampled_schedule(net, 100)
    s2, a2, ap2, b2 = S._subsample(net.alphas, net.alphas_prev, net.betas, 100)
    assert list(steps) == list(range(0, 1000, 10)) and torch.equal(a, a2) and torch.equal(ap, ap2) and torch.equal(b, b2)
    steps, a, ap, b = M._subsampled_schedule(net, 1000)
    assert len(steps) == 1000 and torch.equal(b, net.betas)


def test_plain_unet_state_dict_matches_reference_layout():
    from evcdiff.models.unet import UNet_DDPM, unet_spec
    from oracle import unet_plain as U
    for mode in ("deep", "deeper"):
        cfg = common.make_config(ngf=32, image_size=16)
        cfg.mode = mode
        net = UNet_DDPM(cfg)
        sd = net.state_dict()
        shapes = U.unet_param_shapes(cfg)
        keys = [k for k in sd if k not in ("betas", "alphas", "alphas_prev")]
        assert keys == list(shapes.keys())
        assert all(tuple(sd[k].shape) == tuple(shapes[k]) for k in keys)
        ours, theirs = unet_spec(cfg), U.unet_spec(cfg)
        assert ours["down"] == theirs["down"] and ours["mid"] == theirs["mid"] and ours["up"] == theirs["up"]
        assert float(sd["unet.out.weight"].abs().max()) < 1e-4  # zero-init output conv (unet.py:243)


def test_checkpoint_loading_flow_like_city_sender():
    """city_sender.py:304-324: DataParallel wrap, load `module.`-prefixed state, EMA shadow copy, unwrap."""
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    from evcdiff.models.ema import EMAHelper
    cfg = common.tiny_config()
    src = UNetMore_DDPM(cfg)
    states = [{"module." + k: v.clone() for k, v in src.state_dict().items()},
              {k: torch.full_like(p, 0.25) for k, p in src.named_parameters()}]
    scorenet = torch.nn.DataParallel(UNetMore_DDPM(cfg))
    v_init = scorenet.module.unet._weights_version()
    scorenet.load_state_dict(states[0], strict=False)
    scorenet.eval()
    v0 = scorenet.module.unet._weights_version()  # read AFTER load_state_dict: ema() alone must change it
    assert v0 != v_init
    ema_helper = EMAHelper(mu=cfg.model.ema_rate)
    ema_helper.register(scorenet)
    ema_helper.load_state_dict(states[-1])
    ema_helper.ema(scorenet)
    net = scorenet.module if hasattr(scorenet, "module") else scorenet
    assert all(bool((p == 0.25).all()) for p in net.parameters())
    v1 = net.unet._weights_version()
    assert v1 != v0  # the engine will repack its operands
    assert v1[1] != v0[1]  # ... through the parameters' own version counters, not only the explicit epoch
    with torch.no_grad():
        next(net.parameters()).data.mul_(2.0)  # invisible to version counters
    assert net.unet._weights_version() == v1
    net.unet.mark_weights_changed()
    assert net.unet._weights_version() != v1


def test_pipeline_get_model_loads_checkpoint_once():
    """evcdiff.pipeline.get_model = city_sender.py:304-324 without the per-cycle reload."""
    from evcdiff.models.better.ncsnpp_more import UNetMore_DDPM
    from evcdiff.pipeline import get_model
    cfg = common.tiny_config()
    src = UNetMore_DDPM(cfg)
    states = [{"module." + k: v.clone() for k, v in src.state_dict().items()}, "optimizer", 7,
              {k: torch.full_like(p, -0.5) for k, p in src.named_parameters()}]
    net = get_model(cfg, states=states)
    assert not net.training and all(bool((p == -0.5).all()) for p in net.parameters())
    cfg.model.ema = False
    net = get_model(cfg, states=states)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), src.state_dict().values()))


def test_tile_planning_host_logic():
    """Host-side mirrors of the kernel's tiling (ops.m_tiles / ops.pick_bn) and the engine's fusion rules."""
    from evcdiff import ops
    from evcdiff.engine import EngineBase
    # 128-row M tiles: one image row at 128x128, whole small images batched into a tile at 8x8
    assert ops.m_tiles(46, 128, 128, False) == 46 * 128
    assert ops.m_tiles(46, 64, 64, False) == 46 * 32
    assert ops.m_tiles(46, 8, 8, False) == 23 and ops.m_tiles(5, 8, 8, False) == 3
    assert ops.m_tiles(4, 8, 8, True) == 4  # per-sample B operand: never two samples in a tile
    # large problems take the widest tile that divides N; N = 15 falls back to a padded 16-wide tile
    assert ops.pick_bn(192, ops.m_tiles(46, 128, 128, False), 27) == 192
    assert ops.pick_bn(15) == 16 and ops.pick_bn(768) == 256
    for n in (192, 384, 576, 768, 1152, 1536):
        bn = ops.pick_bn(n, ops.m_tiles(46, 8, 8, False), 108)
        assert n % bn == 0 and bn % 16 == 0
    # more tiles than SMs even with the widest tile (6 videos at 64x64: 192 tiles): a narrower N tile re-reads the A tile
    # in every round, so the widest one wins (profiles/r02_bn_probe.txt); with few tiles (6 videos at 16x16) the tile
    # shrinks so that more SMs get work
    assert ops.pick_bn(192, ops.m_tiles(6, 64, 64, False), 54, sms=148) == 192
    assert ops.pick_bn(384, ops.m_tiles(6, 64, 64, False), 54, sms=148) == 192
    assert ops.pick_bn(576, ops.m_tiles(6, 16, 16, False), 81, sms=148) < 192
    # fused attention: head dims are multiples of 64 (128 between 256 and 384); larger heads split their output columns
    assert ops.attn_supported(1024, 384, 2) and ops.attn_supported(64, 768, 4) and ops.attn_supported(4096, 384, 1)
    assert ops.attn_supported(64, 768, 1) and ops.attn_supported(256, 1024, 2)  # unet.py 'deeper': one head of 768
    assert not ops.attn_supported(100, 384, 2) and not ops.attn_supported(256, 320, 1) and not ops.attn_supported(256, 96, 1)
    # fused statistics need whole 32-row warp slices inside one sample; the fused GroupNorm apply whole 128-row tiles
    assert EngineBase.can_fuse_stats(128, 128) and EngineBase.can_fuse_stats(8, 8) and not EngineBase.can_fuse_stats(2, 2)


def test_header_is_valid_c_and_links_from_a_c_host(tmp_path):
    """include/evcdiff.h is the drop-in boundary for non-Python hosts: it must compile as C99 with warnings as errors,
    link against libevcdiff.so, and agree with the library on struct sizes (examples/c_abi_demo.c, CPU part only)."""
    import shutil
    import subprocess
    from evcdiff import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    _lib.load()
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "c_abi_demo")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", libdir, "-levcdiff",
                    "-Wl,-rpath," + libdir, "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "evcdiff ABI version 1" in out


def test_bench_reference_arm_falls_back_to_port(monkeypatch):
    """Without a reachable copy of the reference the arm times the oracle port and says so."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    monkeypatch.setattr(bench, "REF_ROOTS", [None, "/nonexistent"])
    assert bench.find_reference() is None
    base = bench.cpu_baseline(1, one_thread=False)
    assert base["kind"] == "port" and base["value"] > 0


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's CPU path, timed on the host) prints one JSON line with the contract
    keys; it needs no GPU, so it runs here with a one-evaluation sample."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-evals", "1"], check=True, capture_output=True, text=True).stdout
    line = json.loads(out.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    # "reference" when an unmodified copy of the reference is reachable (baseline/_ref, installed by build() in the
    # build container and shipped to the GPU box), else the oracle port
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "models", "__init__.py")) or \
        os.path.isdir("/root/reference/models") or bool(os.environ.get("EVC_REF"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and line["vs_baseline"] is None


# every conv3x3 / 1x1 launch shape of one NCSN++ evaluation at 128x128, (Cin, Cout, H)
_CONV3 = [(192, 192, 128), (384, 192, 128), (192, 192, 64), (384, 384, 64), (384, 192, 64), (576, 192, 64), (192, 192, 32),
          (192, 384, 32), (384, 384, 32), (576, 576, 32), (576, 384, 32), (768, 384, 32), (960, 384, 32), (384, 384, 16),
          (384, 576, 16), (576, 576, 16), (768, 768, 16), (960, 576, 16), (1152, 576, 16), (1344, 576, 16), (576, 576, 8),
          (576, 768, 8), (768, 768, 8), (1344, 768, 8), (1536, 768, 8)]
_CONV1 = [(384, 1152, 32), (576, 1728, 16), (768, 2304, 8), (384, 384, 32), (576, 576, 16), (768, 768, 8)]  # q|k|v and NIN_3


def test_tile_choice_invariants_for_every_batch_size():
    """ops.pick_tile for every per-GPU batch a sharded run can produce (1..64 videos) and every launch shape of the model:
    the N tile divides N, split-K only where the kernel supports it (32-column chunks, >= 4 K blocks per slice, every
    slice resident, partial tiles inside the workspace), and the fused-GroupNorm guard mirrors its two-round rule."""
    from evcdiff import ops
    sms = 148
    for B in range(1, 65):
        for taps, shapes in ((9, _CONV3), (1, _CONV1)):
            for cin, cout, h in shapes:
                mt = ops.m_tiles(B, h, h, False)
                assert mt == -(-B * h * h // 128) or h < 16  # whole images share a tile only below 128 pixels per image
                kblocks = taps * cin // 64
                bn, S = ops.pick_tile(cout, mt, kblocks, sms=sms)
                assert cout % bn == 0 and bn % 16 == 0 and bn <= 256, (B, cin, cout, h, bn)
                assert S >= 1
                if S > 1:
                    tiles = mt * (cout // bn)
                    assert bn % 32 == 0 and kblocks // S >= 4 and tiles * S <= sms, (B, cin, cout, h, bn, S)
                    assert S * ((mt + 1) // 2 * 2) * 128 * cout * 4 <= ops.SPLIT_K_WS_BYTES
                    assert mt * (cout // 256 if cout % 256 == 0 else 1) < sms  # never for launches that fill the GPU
                # the unsplit choice is what pick_bn alone gives
                if S == 1:
                    assert bn == ops.pick_bn(cout, mt, kblocks, sms)
    # fused GroupNorm apply: a sample's tiles must fit in two rounds of the persistent grid (gemm_tc.cu plan guard)
    assert ops.gn_fuse_fits(128, 46 * 128, 1, sms=sms) and ops.gn_fuse_fits(128, 128, 1, sms=sms)
    assert not ops.gn_fuse_fits(512, 512, 1, sms=sms)      # image_size 256: 512 tiles per sample > 2 x 148
    assert not ops.gn_fuse_fits(128, 128, 1, sms=40)       # a reduced SM count breaks the rule as well
