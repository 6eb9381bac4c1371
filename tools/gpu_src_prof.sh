#!/usr/bin/env bash
O=gpurun_out/src
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $O/cur_64 python tools/prof_one.py ours 64 4096 7168 > $O/prof.log 2>&1; echo "rc=$?"
ncu -i $O/cur_64.ncu-rep --page source --csv > $O/cur_64.source.csv 2>/dev/null
ncu -i $O/cur_64.ncu-rep --page raw --csv > $O/cur_64.raw.csv 2>/dev/null
rm -f $O/cur_64.ncu-rep
ls -la $O
