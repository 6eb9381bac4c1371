#!/usr/bin/env bash
# Round-2 GPU pass C (one B200): tests, sweeps after the epilogue / pair-split changes, bench line, sanitizers.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu_c.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_c.log
tail -c 1500 $OUT/pytest_gpu_c.log
timeout 600 python tools/tune.py store > $OUT/tune_store_c.log 2>&1
timeout 400 python tools/tune.py mid2 > $OUT/tune_mid_c.log 2>&1
timeout 300 python tools/tune.py small > $OUT/tune_small_c.log 2>&1
timeout 120 python tools/stamps.py --cold > $OUT/stamps_small_c.log 2>&1
( time timeout 900 python bench.py ) > $OUT/bench_c.log 2> $OUT/bench_c.err
echo "bench rc=$?" >> $OUT/bench_c.err
bash tools/gpu_sanitize.sh
