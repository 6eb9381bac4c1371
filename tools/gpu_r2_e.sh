#!/usr/bin/env bash
# Round-2 GPU pass E (one B200): tests after the contiguous padding skip + transposed staged epilogue, sweeps, bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu_e.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_e.log
tail -c 2000 $OUT/pytest_gpu_e.log
timeout 600 python tools/tune.py swap2 > $OUT/tune_swap2_e.log 2>&1
( time timeout 900 python bench.py ) > $OUT/bench_e.log 2> $OUT/bench_e.err
echo "bench rc=$?" >> $OUT/bench_e.err
for mm in 64 128 256; do timeout 300 python bench.py --workload contiguous --mean-m $mm --steps 10 --warmup 3 > $OUT/bench_contig_m$mm.log 2>&1; done
