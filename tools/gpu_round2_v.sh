#!/bin/bash
# same-box A/B of the PDL trigger position / weight fetch ahead of the PDL wait (builds in build/, see profiles/r02_notes.md)
mkdir -p gpurun_out
bash tools/gpu_ab_lib.sh r2v build/libevcdiff_nopre.so build/libevcdiff_early.so build/libevcdiff_earlypre.so
for l in early earlypre; do
  EVC_PDL=1 EVC_LIB=$PWD/build/libevcdiff_$l.so python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos 46 > gpurun_out/r2v_pdl1_${l}_b46.json 2> gpurun_out/r2v_pdl1_${l}_b46.err
  python -c "
import json; d=json.load(open('gpurun_out/r2v_pdl1_${l}_b46.json')); print('$l EVC_PDL=1 B=46', round(d['value'],2), 'clk', d['clocks']['sm_mhz'])"
  EVC_LIB=$PWD/build/libevcdiff_$l.so python bench.py --steps 2 --warmup 1 --no-cpu-baseline --videos 1 > gpurun_out/r2v_${l}_b1.json 2> gpurun_out/r2v_${l}_b1.err
  python -c "
import json; d=json.load(open('gpurun_out/r2v_${l}_b1.json')); print('$l B=1', round(d['value'],2), 'clk', d['clocks']['sm_mhz'])"
done
for l in nopre early earlypre; do echo == $l; EVC_LIB=$PWD/build/libevcdiff_$l.so python tools/gpu_launch_floor.py 2>&1 | grep "pdl=1" | tee gpurun_out/r2v_floor_$l.txt; done
