#!/bin/bash
# round 2, call D: whole GPU suite with split-K in, then bench at 46 / 6 / 1 videos with per-launch profiles
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rf --durations=8 2>&1 | tail -60 > gpurun_out/r2d_pytest.txt; tail -30 gpurun_out/r2d_pytest.txt
for b in 46 6 1; do
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2d_prof_b$b.json > gpurun_out/r2d_bench_b$b.json 2> gpurun_out/r2d_bench_b$b.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2d_bench_b$b.json'))
    print('B=$b', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'gemm all', round(d['roofline']['all_gemm_launches']['frac'],3), 'clk', d['clocks']['sm_mhz'], 'launches', d['gpu_launches'])
except Exception as e:
    print('B=$b failed', e, open('gpurun_out/r2d_bench_b$b.err').read()[-1500:])
PY
done
