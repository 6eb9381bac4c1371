#!/usr/bin/env bash
# Round-2 GPU pass K (one B200): transposed-output kernel with generic coalesced stores out of the per-warp staging buffers.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider ) > $OUT/pytest_gpu_k.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_k.log
tail -c 800 $OUT/pytest_gpu_k.log
timeout 900 python tools/tune.py swap3 > $OUT/tune_k_swap3.log 2>&1
timeout 300 python tools/tune.py store_exp > $OUT/tune_k_store_base.log 2>&1
DGB200_LIB=$PWD/deepgemm_b200/lib/libdgb200_nostore.so timeout 300 python tools/tune.py store_exp > $OUT/tune_k_store_nostore.log 2>&1
