#!/bin/bash
# round 2, call A: does the warp-uniform issue path hold (tests), what does it buy (bench), microbenchmark + cuBLAS yardstick
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
./build/umma_microbench 2048 > gpurun_out/umma_microbench.jsonl 2> gpurun_out/umma_microbench.err; echo "microbench rc=$?"; tail -40 gpurun_out/umma_microbench.jsonl
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for b in 46 6 1; do
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos $b --micro-batch $b > gpurun_out/r2a_bench_b$b.json 2> gpurun_out/r2a_bench_b$b.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2a_bench_b$b.json'))
    print('B=$b', round(d['value'],2), 'fps', d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'gemm all', round(d['roofline']['all_gemm_launches']['frac'],3), 'dom', round(d['roofline']['frac'],3), 'clk', d['clocks'])
except Exception as e:
    print('B=$b failed', e, open('gpurun_out/r2a_bench_b$b.err').read()[-800:])
PY
done
python tools/gpu_cublas_yardstick.py gpurun_out/cublas_yardstick.json 2>&1 | tail -8
