#!/bin/bash
# final records of the round: whole GPU suite + smoke, the default bench line (with the CPU baseline and the GPU-eager
# context), launch list + ncu --set full captures of the hot kernels
mkdir -p gpurun_out
bash tools/gpu_tests.sh tests/test_gemm_gpu.py tests/test_gemm_pair_gpu.py tests/test_elementwise_gpu.py tests/test_attn_gpu.py tests/test_ncsnpp_gpu.py tests/test_full_model_gpu.py tests/test_fp32_mode_gpu.py tests/test_sender_gpu.py tests/test_unet_plain_gpu.py tests/test_two_gpu.py
python bench.py --gpu-eager-context > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; tail -c 600 gpurun_out/r2f_bench_default.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2f_bench_default.json'))
print('default bench', round(d['value'],2), 'fps e2e', round(d['e2e']['value'],2), d['roofline']['ms_per_eval_by_kernel'], 'step_frac', round(d['roofline']['step_tensor_frac'],3), 'clk', d['clocks'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline'].get('value_1thread'))
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; cut -c1-300 gpurun_out/r2f_bench_reference.json
bash tools/gpu_profile.sh 2>&1 | tail -12
