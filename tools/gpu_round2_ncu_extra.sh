#!/bin/bash
# ncu --set full of the two kernels that had only CUDA-event numbers so far: the vectorised sampler update
# (state_update_vec_kernel, 46 videos) and the fused attention kernel at the unet.py 'deep' shape (N = 4096 keys, d = 384).
# Same command plain first (must exit 0), then under ncu; reports exported to CSV on the box.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
CMD1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
CMD2="python bench.py --model unet_deep --videos 32 --steps 1 --warmup 1 --no-cpu-baseline --no-profile"
$CMD1 > gpurun_out/x_plain1.log 2> gpurun_out/x_plain1.err || { echo "plain run 1 failed"; tail -5 gpurun_out/x_plain1.err; exit 1; }
timeout 200 ncu --set full --clock-control none --import-source on -k regex:state_update_vec_kernel -s 2 -c 3 -o gpurun_out/x_upd $CMD1 > gpurun_out/x_ncu1.log 2>&1
echo "sampler-update capture rc=$?"
ncu -i gpurun_out/x_upd.ncu-rep --page raw --csv > gpurun_out/x_ncu_full_sampler_update.csv 2>/dev/null
ncu -i gpurun_out/x_upd.ncu-rep --page details --csv 2>/dev/null | grep -iE "stall|warp cycles|Eligible|Achieved Occupancy|Registers|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth" | head -60 > gpurun_out/x_ncu_details_sampler_update.csv
rm -f gpurun_out/x_upd.ncu-rep
$CMD2 > gpurun_out/x_plain2.log 2> gpurun_out/x_plain2.err || { echo "plain run 2 failed"; tail -5 gpurun_out/x_plain2.err; exit 1; }
timeout 240 ncu --set full --clock-control none --import-source on -k regex:evc_attn_kernel -s 1 -c 4 -o gpurun_out/x_attn $CMD2 > gpurun_out/x_ncu2.log 2>&1
echo "attention capture rc=$?"
ncu -i gpurun_out/x_attn.ncu-rep --page raw --csv > gpurun_out/x_ncu_full_attn_deep.csv 2>/dev/null
rm -f gpurun_out/x_attn.ncu-rep
python tools/ncu_table.py gpurun_out/x_ncu_full_sampler_update.csv "state_update_vec_kernel, 46 videos" | cut -c1-220
python tools/ncu_table.py gpurun_out/x_ncu_full_attn_deep.csv "evc_attn_kernel, unet.py deep, 32 videos" | cut -c1-220
tail -2 gpurun_out/x_plain1.log | cut -c1-200; tail -2 gpurun_out/x_plain2.log | cut -c1-200
