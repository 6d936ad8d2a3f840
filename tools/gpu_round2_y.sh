#!/bin/bash
# ncu --set full of the fused res-block prologue kernels (gn_fir down / up) at the benchmark shapes + warp stall reasons
mkdir -p gpurun_out
python tools/gpu_gn_fir_bench.py > gpurun_out/r2y_plain.log 2>&1 || { tail -5 gpurun_out/r2y_plain.log; exit 1; }
for k in gn_fir_down_kernel gn_fir_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/r2y_$k python tools/gpu_gn_fir_bench.py > gpurun_out/r2y_ncu_$k.log 2>&1
  ncu -i gpurun_out/r2y_$k.ncu-rep --page raw --csv > gpurun_out/r2y_$k.csv 2>/dev/null
  ncu -i gpurun_out/r2y_$k.ncu-rep --page source --csv > gpurun_out/r2y_${k}_source.csv 2>/dev/null
  rm -f gpurun_out/r2y_$k.ncu-rep
  python tools/ncu_summary.py gpurun_out/r2y_$k.csv
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r2y_$k.csv')))
hdr,units,d=rows[0],rows[1],rows[2]
st=[(float(d[i].replace(',','')),h) for i,h in enumerate(hdr) if 'issue_stalled' in h and 'per_issue_active' in h and d[i] not in ('','n/a')]
for v,h in sorted(st,reverse=True)[:8]: print(round(v,3),h)
PY
done
