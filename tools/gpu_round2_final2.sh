#!/bin/bash
# last records of round 2 (after the bench label / roofline-denominator change and the sender's LPIPS rule): whole GPU suite +
# smoke, the default bench line exactly as the driver runs it, the reference arm
mkdir -p gpurun_out
bash tools/gpu_tests.sh tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_elementwise_gpu.py tests/test_attn_gpu.py tests/test_ncsnpp_gpu.py tests/test_full_model_gpu.py tests/test_fp32_mode_gpu.py tests/test_sender_gpu.py tests/test_unet_plain_gpu.py tests/test_two_gpu.py
cp gpurun_out/tests.log gpurun_out/r2x_gpu_tests.log
python bench.py > gpurun_out/r2x_bench_default.json 2> gpurun_out/r2x_bench_default.err; tail -c 600 gpurun_out/r2x_bench_default.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2x_bench_default.json'))
r=d['roofline']
print('default bench', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), r['ms_per_eval_by_kernel'], 'frac', round(r['frac'],3), 'all', round(r['all_gemm_launches']['frac'],3), 'step_frac', round(r['step_tensor_frac'],3), 'clk', d['clocks'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('value_1thread'))
print(d['config']['workload'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err; cut -c1-300 gpurun_out/r2x_bench_reference.json
