#!/bin/bash
# ncu evidence for profiles/: launch list (device time per launch) + one full capture of the dominant kernel.
# Same command line plain first (must exit 0), then under ncu.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2> gpurun_out/plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:evc_gemm_kernel -s 1 -c 3 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
tail -3 gpurun_out/ncu_launches.log gpurun_out/ncu_full.log
