#!/bin/bash
# Builds libevcdiff.so with the GEMM wait-time probe compiled in (experiments only; `python __graft_entry__.py build`
# restores the shipped library).
P=extreme-video-compression-with-prediction-using-pre-trainded-diffusion-models-_b200
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --cudart static -shared -DEVC_GEMM_PROF \
  -o $P/evcdiff/lib/libevcdiff.so $P/csrc/evc_host.cu $P/csrc/gemm_tc.cu $P/csrc/attn_tc.cu $P/csrc/elementwise.cu $P/csrc/sampler.cu
