#!/usr/bin/env bash
# Round-2 GPU pass H (one B200): RED accumulate epilogue, BF16 split-K tests, param-warm A/B, swapped-kernel profile.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_bf16_gpu.py tests/test_gemm_gpu.py tests/test_round2_gpu.py -m gpu -q -p no:cacheprovider ) > $OUT/pytest_gpu_h.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_h.log
tail -c 1500 $OUT/pytest_gpu_h.log
timeout 300 python tools/bf16_bench.py > $OUT/bf16_bench_h.log 2>&1
timeout 200 python tools/kgrouped_bench.py > $OUT/kgrouped_bench_h.log 2>&1
SH=64x4096x7168,64x7168x2048,128x7168x2048,512x4096x7168
timeout 200 python tools/stamps.py --cold --shapes=$SH > $OUT/stamps_h_base.log 2>&1
DGB200_LIB=$PWD/deepgemm_b200/lib/libdgb200_warm.so timeout 200 python tools/stamps.py --cold --shapes=$SH > $OUT/stamps_h_warm.log 2>&1
timeout 300 python tools/tune.py ab_small > $OUT/tune_h_base.log 2>&1
DGB200_LIB=$PWD/deepgemm_b200/lib/libdgb200_warm.so timeout 300 python tools/tune.py ab_small > $OUT/tune_h_warm.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
DGB200_SWAP=1 DGB200_BLOCK_M=224 DGB200_TMA_STORE=1 timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $OUT/ours_swap_k2048 python tools/prof_r2.py dense ours 4096 7168 2048 > $OUT/prof_swap.log 2>&1
ncu -i $OUT/ours_swap_k2048.ncu-rep --page raw --csv > $OUT/ours_swap_k2048.raw.csv 2>/dev/null
ncu -i $OUT/ours_swap_k2048.ncu-rep --page source --csv > $OUT/ours_swap_k2048.source.csv 2>/dev/null
rm -f $OUT/ours_swap_k2048.ncu-rep
