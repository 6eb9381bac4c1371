#!/usr/bin/env bash
# Round-2 GPU pass L (one B200): full GPU suite on the current build, k-grouped sweeps.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu_l.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_l.log
tail -c 1200 $OUT/pytest_gpu_l.log
timeout 600 python tools/kgrouped_bench.py > $OUT/kgrouped_bench_l.log 2>&1
