#!/bin/bash
# round 2, call E: split-K with the coalesced workspace layout (tests + bench at 6 / 46 videos), then the probe build:
# operand-skip experiments and the fused-GroupNorm epilogue
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -rf -k split_k 2>&1 | tail -8
for b in 6 46; do
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --profile-json gpurun_out/r2e_prof_b$b.json > gpurun_out/r2e_bench_b$b.json 2> gpurun_out/r2e_bench_b$b.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2e_bench_b$b.json'))
    print('B=$b', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks']['sm_mhz'])
except Exception as e:
    print('B=$b failed', e, open('gpurun_out/r2e_bench_b$b.err').read()[-1500:])
PY
done
EVC_GEMM_SPLIT_K=0 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos 6 > gpurun_out/r2e_bench_b6_nosplit.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2e_bench_b6_nosplit.json')); print('B=6 no split', round(d['value'],2), d['roofline']['ms_per_eval_by_kernel'])"
export EVC_LIB=$PWD/build/libevcdiff_prof.so
SKIP=1 python tools/gpu_gemm_waits.py 2>&1 | cut -c1-150 | tee gpurun_out/r2e_skip.txt
GNFUSE=1 python tools/gpu_gemm_waits.py 2>&1 | tee gpurun_out/r2e_gnfuse.txt
