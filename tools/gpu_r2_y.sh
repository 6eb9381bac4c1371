#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2p; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $O/ours_4096 python tools/prof_r2.py dense ours 4096 4096 7168 > $O/prof_ours_4096.log 2>&1; echo "rc=$?"
ncu -i $O/ours_4096.ncu-rep --page raw --csv > $O/ours_4096.raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:sm100_fp8 -s 1 -c 1 -f -o $O/ref_4096 python tools/prof_r2.py dense ref 4096 4096 7168 > $O/prof_ref_4096.log 2>&1; echo "rc=$?"
ncu -i $O/ref_4096.ncu-rep --page raw --csv > $O/ref_4096.raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
