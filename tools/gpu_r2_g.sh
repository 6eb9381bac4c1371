#!/usr/bin/env bash
# Round-2 GPU pass G (one B200): BF16 fixes (accumulate epilogue, cluster split-K), param-warm experiment, k-grouped A/B.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_bf16_gpu.py tests/test_gemm_gpu.py tests/test_round2_gpu.py -m gpu -q -x -p no:cacheprovider ) > $OUT/pytest_gpu_g.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_g.log
tail -c 1500 $OUT/pytest_gpu_g.log
timeout 300 python tools/bf16_bench.py > $OUT/bf16_bench_g.log 2>&1
timeout 200 python tools/kgrouped_bench.py > $OUT/kgrouped_bench_g.log 2>&1
SH=64x4096x7168,64x7168x2048,128x7168x2048,512x4096x7168
timeout 200 python tools/stamps.py --cold --shapes=$SH > $OUT/stamps_g_base.log 2>&1
DGB200_LIB=$PWD/deepgemm_b200/lib/libdgb200_warm.so timeout 200 python tools/stamps.py --cold --shapes=$SH > $OUT/stamps_g_warm.log 2>&1
timeout 300 python tools/tune.py ab_small > $OUT/tune_g_base.log 2>&1
DGB200_LIB=$PWD/deepgemm_b200/lib/libdgb200_warm.so timeout 300 python tools/tune.py ab_small > $OUT/tune_g_warm.log 2>&1
