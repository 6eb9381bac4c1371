#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
( time timeout 900 python -m pytest tests/test_bf16_gpu.py tests/test_round2_gpu.py -m gpu -q -x -p no:cacheprovider ) > $OUT/pytest_gpu_q.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_q.log; tail -c 1500 $OUT/pytest_gpu_q.log
timeout 600 python tools/bf16_bench.py > $OUT/bf16_bench_q.log 2>&1; grep bmk $OUT/bf16_bench_q.log
timeout 200 python tools/host_overhead.py > $OUT/host_overhead_q.log 2>&1; tail -1 $OUT/host_overhead_q.log
