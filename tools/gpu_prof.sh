#!/usr/bin/env bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o gpurun_out/ours_4096 python tools/prof_one.py ours 4096 4096 7168 > gpurun_out/prof_ours_4096.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:sm100_fp8 -s 1 -c 1 -f -o gpurun_out/ref_4096 python tools/prof_one.py ref 4096 4096 7168 > gpurun_out/prof_ref_4096.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o gpurun_out/ours_k2048 python tools/prof_one.py ours 4096 7168 2048 > gpurun_out/prof_ours_k2048.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:sm100_fp8 -s 1 -c 1 -f -o gpurun_out/ref_k2048 python tools/prof_one.py ref 4096 7168 2048 > gpurun_out/prof_ref_k2048.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o gpurun_out/ours_64 python tools/prof_one.py ours 64 4096 7168 > gpurun_out/prof_ours_64.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:sm100_fp8 -s 1 -c 1 -f -o gpurun_out/ref_64 python tools/prof_one.py ref 64 4096 7168 > gpurun_out/prof_ref_64.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
