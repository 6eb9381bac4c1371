#!/usr/bin/env bash
# Round-2 GPU pass D (one B200): second orientation tests + sweep, gap probe, sanitizers, host-path timing.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_round2_gpu.py tests/test_gemm_gpu.py -m gpu -q -p no:cacheprovider ) > $OUT/pytest_gpu_d.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_d.log
tail -c 2500 $OUT/pytest_gpu_d.log
timeout 900 python tools/tune.py swap > $OUT/tune_swap.log 2>&1
timeout 300 python tools/gap_probe.py > $OUT/gap_probe.log 2>&1
bash tools/gpu_sanitize.sh
