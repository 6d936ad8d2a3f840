#!/bin/bash
# ncu evidence for profiles/: launch list (device time per launch) + full captures of the hot kernels of one UNet
# evaluation.  Same command line plain first (must exit 0), then under ncu.  The .ncu-rep files are exported to CSV on
# the box and deleted (gpurun only copies back 64 MiB).
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
K="regex:evc_gemm_kernel|gn_apply_kernel|gn_apply_stream_kernel|gn_fir_kernel|gn_fir_down_kernel|evc_attn_kernel"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 520 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
# 36 + 44 hot kernels from the eager warm-up evaluations (both windows together cover every launch shape of the
# 128x128 and 64x64 levels, including the dominant K=3456 convolution, gn_fir up and down, and the output conv)
ncu --set full --clock-control none -k "$K" -s 0 -c 60 -o gpurun_out/prof_head $CMD > gpurun_out/ncu_head.log 2>&1
echo "head capture rc=$?"
ncu -i gpurun_out/prof_head.ncu-rep --page raw --csv > gpurun_out/ncu_full_head.csv 2>/dev/null; rm -f gpurun_out/prof_head.ncu-rep
ncu --set full --clock-control none -k "$K" -s 110 -c 60 -o gpurun_out/prof_tail $CMD > gpurun_out/ncu_tail.log 2>&1
echo "tail capture rc=$?"
ncu -i gpurun_out/prof_tail.ncu-rep --page raw --csv > gpurun_out/ncu_full_tail.csv 2>/dev/null; rm -f gpurun_out/prof_tail.ncu-rep
ls -la gpurun_out/ | head -30; du -sh gpurun_out
