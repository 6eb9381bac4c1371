#!/usr/bin/env bash
# Round-2 GPU pass I (one B200): one-tile-per-cluster grid A/B, swapped-epilogue experiments.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_round2_gpu.py tests/test_bf16_gpu.py -m gpu -q -x -p no:cacheprovider ) > $OUT/pytest_gpu_i.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_i.log
tail -c 800 $OUT/pytest_gpu_i.log
SH=64x7168x2048,128x7168x2048,512x4096x7168
timeout 200 python tools/stamps.py --cold --shapes=$SH > $OUT/stamps_i_grid.log 2>&1
DGB200_GRID_TILES=0 timeout 200 python tools/stamps.py --cold --shapes=$SH > $OUT/stamps_i_nogrid.log 2>&1
timeout 300 python tools/tune.py ab_small > $OUT/tune_i_grid.log 2>&1
DGB200_GRID_TILES=0 timeout 300 python tools/tune.py ab_small > $OUT/tune_i_nogrid.log 2>&1
timeout 300 python tools/tune.py swap_exp > $OUT/tune_i_swap_base.log 2>&1
DGB200_LIB=$PWD/deepgemm_b200/lib/libdgb200_nostore.so timeout 300 python tools/tune.py swap_exp > $OUT/tune_i_swap_nostore.log 2>&1
DGB200_LIB=$PWD/deepgemm_b200/lib/libdgb200_nostage.so timeout 300 python tools/tune.py swap_exp > $OUT/tune_i_swap_nostage.log 2>&1
