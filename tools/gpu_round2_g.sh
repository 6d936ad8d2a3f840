#!/bin/bash
# A/B of the number of epilogue warps (8 / 12 / 16) on one box: GEMM tests, then bench at 46 and 6 videos
mkdir -p gpurun_out
for w in 8 12 16; do
  export EVC_LIB=$PWD/build/libevcdiff_w$w.so
  timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py -m gpu -q -x 2>&1 | tail -2
  for b in 46 6; do
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2g_prof_w${w}_b$b.json > gpurun_out/r2g_bench_w${w}_b$b.json 2> gpurun_out/r2g_bench_w${w}_b$b.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2g_bench_w${w}_b$b.json'))
    print('warps $w B=$b', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('warps $w B=$b failed', e, open('gpurun_out/r2g_bench_w${w}_b$b.err').read()[-1500:])
PY
  done
done
