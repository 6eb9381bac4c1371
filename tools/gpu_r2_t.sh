#!/usr/bin/env bash
# ncu --set full captures of the kernels that changed since the profiles pass (contiguous padding skip, masked, dominant dense, small M)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2p; mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 rx=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$rx -s $skip -c 1 -f -o $O/$name "$@" > $O/prof_$name.log 2>&1; echo "$name rc=$?"; }
prof ours_4096 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 4096 4096 7168
prof ours_64 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 64 4096 7168
prof ours_contig fp8_gemm_kernel 1 python tools/prof_r2.py contiguous ours 128
prof ref_contig sm100_fp8 1 python tools/prof_r2.py contiguous ref 128
prof ours_masked fp8_gemm_kernel 1 python tools/prof_r2.py masked ours 64
for n in ours_4096 ours_64 ours_contig ref_contig ours_masked; do ncu -i $O/$n.ncu-rep --page raw --csv > $O/$n.raw.csv 2>/dev/null; done
ncu -i $O/ours_contig.ncu-rep --page source --csv > $O/ours_contig.source.csv 2>/dev/null
find $O -name "*.ncu-rep" -delete
ls -la $O | head -40
