#!/usr/bin/env bash
# EP tests on one GPU after the combine changes
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
( time timeout 600 python -m pytest tests/test_ep_gpu.py -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu_m.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_m.log
tail -c 600 $OUT/pytest_gpu_m.log
