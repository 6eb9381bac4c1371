#!/usr/bin/env bash
# Round-2 GPU pass F (one B200): BF16-operand family tests + reference digests, full suite, bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
( time timeout 600 python tests/golden/make_golden_digests.py --only bf16 --parts 3 ) > $OUT/digests_bf16.log 2>&1
cp gpurun_out/gpu_digests.json tests/golden/gpu_digests.json 2>/dev/null
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu_f.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_f.log
tail -c 3000 $OUT/pytest_gpu_f.log
timeout 300 python tools/bf16_bench.py > $OUT/bf16_bench.log 2>&1
( time timeout 900 python bench.py ) > $OUT/bench_f.log 2> $OUT/bench_f.err
echo "bench rc=$?" >> $OUT/bench_f.err
timeout 200 python tools/stamps.py --cold --shapes=64x7168x2048,128x24576x1536,128x7168x2048 > $OUT/stamps_shortk.log 2>&1
timeout 400 python tools/tune.py small3 > $OUT/tune_small3.log 2>&1
