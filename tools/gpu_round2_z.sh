#!/bin/bash
# same-box A/B of the N-tile cost model (EVC_PICK_MODEL=old|new) at 46 / 6 / 5 videos
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_ncsnpp_gpu.py tests/test_full_model_gpu.py -m gpu -q -x --tb=short 2>&1 | tail -5
for m in old new; do
  for b in 46 6 5; do
    EVC_PICK_MODEL=$m python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2z_prof_${m}_b$b.json > gpurun_out/r2z_${m}_b$b.json 2> gpurun_out/r2z_${m}_b$b.err
    python -c "
import json; d=json.load(open('gpurun_out/r2z_${m}_b$b.json')); print('EVC_PICK_MODEL=$m B=$b', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'clk', d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2z_${m}_b$b.err
  done
done
