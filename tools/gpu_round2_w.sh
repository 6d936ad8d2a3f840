#!/bin/bash
# one q|k|v projection + V rows (MN-major) in the attention kernel vs the q|k + V^T layout; unet.py 'deeper' with the
# fused kernel for its 768-wide head
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_ncsnpp_gpu.py tests/test_full_model_gpu.py tests/test_unet_plain_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -8
for q in 0 1; do
  for b in 46 6; do
    EVC_FUSED_QKV=$q python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2w_prof_qkv${q}_b$b.json > gpurun_out/r2w_qkv${q}_b$b.json 2> gpurun_out/r2w_qkv${q}_b$b.err
    python -c "
import json; d=json.load(open('gpurun_out/r2w_qkv${q}_b$b.json')); print('EVC_FUSED_QKV=$q B=$b', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'launches', d['gpu_launches'], 'clk', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2w_qkv${q}_b$b.err
  done
  for m in deep deeper; do
    EVC_FUSED_QKV=$q python bench.py --steps 2 --warmup 1 --no-cpu-baseline --model unet_$m --videos 32 --micro-batch 32 > gpurun_out/r2w_qkv${q}_unet_$m.json 2> gpurun_out/r2w_qkv${q}_unet_$m.err
    python -c "
import json; d=json.load(open('gpurun_out/r2w_qkv${q}_unet_$m.json')); print('EVC_FUSED_QKV=$q unet_$m', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'clk', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2w_qkv${q}_unet_$m.err
  done
done
