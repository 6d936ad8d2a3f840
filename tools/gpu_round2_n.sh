#!/bin/bash
# cooperative split-K + sender parity tests; bench with split-K on / off on one box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_sender_gpu.py tests/test_ncsnpp_gpu.py tests/test_full_model_gpu.py -m gpu -q -x 2>&1 | tail -5
for sk in 1 0; do
for b in 6 1 46; do
  EVC_GEMM_SPLIT_K=$sk python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2n_prof_sk${sk}_b$b.json > gpurun_out/r2n_bench_sk${sk}_b$b.json 2> gpurun_out/r2n_bench_sk${sk}_b$b.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2n_bench_sk${sk}_b$b.json'))
    print('splitk=$sk B=$b', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('splitk=$sk B=$b failed', e, open('gpurun_out/r2n_bench_sk${sk}_b$b.err').read()[-1500:])
PY
done
done
