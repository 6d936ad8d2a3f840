#!/bin/bash
# ncu evidence for profiles/: launch list (device time per launch) + full captures of the dominant kernels.
# Same command line plain first (must exit 0), then under ncu.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:evc_gemm_kernel -s 1 -c 2 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "gemm capture rc=$?"
ncu --set full --clock-control none -k regex:evc_gemm_kernel -s 88 -c 14 -o gpurun_out/prof_gemm_up $CMD > gpurun_out/ncu_gemm2.log 2>&1
echo "gemm (up path) capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gn_apply_kernel -s 2 -c 2 -o gpurun_out/prof_gn_apply $CMD > gpurun_out/ncu_gn.log 2>&1
echo "gn_apply capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:evc_attn_kernel -c 2 -o gpurun_out/prof_attn $CMD > gpurun_out/ncu_attn.log 2>&1
echo "attn capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fir_up_kernel -s 6 -c 2 -o gpurun_out/prof_fir_up $CMD > gpurun_out/ncu_fir.log 2>&1
echo "fir capture rc=$?"
ls -la gpurun_out/ | head -30
