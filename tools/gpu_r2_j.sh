#!/usr/bin/env bash
# Round-2 GPU pass J (one B200): 64-column staged stores of the transposed-output kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -p no:cacheprovider ) > $OUT/pytest_gpu_j.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_j.log
tail -c 800 $OUT/pytest_gpu_j.log
timeout 900 python tools/tune.py swap3 > $OUT/tune_j_swap3.log 2>&1
