#!/bin/bash
# same-box A/B of library builds: tools/gpu_ab_lib.sh <tag> <libA> <libB> ... ; runs the GEMM + model tests on the last one
mkdir -p gpurun_out
tag=$1; shift
last="${@: -1}"
EVC_LIB=$PWD/$last timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_ncsnpp_gpu.py tests/test_full_model_gpu.py -m gpu -q -x 2>&1 | tail -3
for lib in "$@"; do
  name=$(basename $lib .so)
  for b in 46 6; do
    EVC_LIB=$PWD/$lib python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/${tag}_prof_${name}_b$b.json > gpurun_out/${tag}_bench_${name}_b$b.json 2> gpurun_out/${tag}_bench_${name}_b$b.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}_bench_${name}_b$b.json'))
    print('$name B=$b', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('$name B=$b failed', e, open('gpurun_out/${tag}_bench_${name}_b$b.err').read()[-1500:])
PY
  done
done
