#!/usr/bin/env bash
# Round-2 GPU pass A (one B200): full GPU test-suite, golden digests of the reference kernel, the default bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" >> $OUT/smoke.log
( time timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
( time timeout 900 python tests/golden/make_golden_digests.py ) > $OUT/digests.log 2>&1
echo "digests rc=$?" >> $OUT/digests.log
( time timeout 900 python bench.py ) > $OUT/bench.log 2> $OUT/bench.err
echo "bench rc=$?" >> $OUT/bench.err
tail -c 3000 $OUT/pytest_gpu.log
tail -c 1500 $OUT/bench.err
