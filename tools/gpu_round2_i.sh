#!/bin/bash
# coordinator warp for the fused GroupNorm apply + bias staged once: tests, then bench at 46 / 6 videos
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 1200 python -m pytest tests -m gpu -q -rf --deselect tests/test_gemm_gpu.py --deselect tests/test_gemm_pair_gpu.py 2>&1 | tail -12
for b in 46 6; do
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2i_prof_b$b.json > gpurun_out/r2i_bench_b$b.json 2> gpurun_out/r2i_bench_b$b.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2i_bench_b$b.json'))
    print('B=$b', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('B=$b failed', e, open('gpurun_out/r2i_bench_b$b.err').read()[-1500:])
PY
done
