#!/usr/bin/env bash
# Round-2 GPU pass B (one B200): remaining tests, epilogue / mid-M sweeps, stamps, UMMA-N probe.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu_b.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_b.log
timeout 120 python tools/peak_sweep.py > $OUT/peak_sweep.log 2>&1
timeout 600 python tools/tune.py store > $OUT/tune_store.log 2>&1
timeout 600 python tools/tune.py mid > $OUT/tune_mid.log 2>&1
timeout 120 python tools/stamps.py --cold > $OUT/stamps_small.log 2>&1
timeout 120 python tools/stamps.py --cold --big > $OUT/stamps_big.log 2>&1
tail -c 2500 $OUT/pytest_gpu_b.log
