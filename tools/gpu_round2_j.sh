#!/bin/bash
# same box: automatic CTA grouping vs pairs wherever possible, 46 and 6 videos
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_ncsnpp_gpu.py -m gpu -q -x 2>&1 | tail -3
for cgv in 0 2; do
for b in 46 6; do
  EVC_GEMM_CTA_GROUP=$cgv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2j_prof_cg${cgv}_b$b.json > gpurun_out/r2j_bench_cg${cgv}_b$b.json 2> gpurun_out/r2j_bench_cg${cgv}_b$b.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2j_bench_cg${cgv}_b$b.json'))
    print('cg=$cgv B=$b', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('cg=$cgv B=$b failed', e, open('gpurun_out/r2j_bench_cg${cgv}_b$b.err').read()[-1500:])
PY
done
done
