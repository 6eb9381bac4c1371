#!/usr/bin/env bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o gpurun_out/ours_contig python tools/prof_grouped.py ours 48 256 > gpurun_out/prof_ours_contig.log 2>&1; echo "rc=$?"
timeout 900 $NCU -k regex:sm100_fp8 -s 1 -c 1 -f -o gpurun_out/ref_contig python tools/prof_grouped.py ref 48 256 > gpurun_out/prof_ref_contig.log 2>&1; echo "rc=$?"
tail -2 gpurun_out/prof_ours_contig.log gpurun_out/prof_ref_contig.log
