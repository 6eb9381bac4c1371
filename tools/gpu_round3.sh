#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/bringup.py ref > gpurun_out/ref.log 2>&1; echo "ref rc=$?"
grep ref_vs_ours gpurun_out/ref.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['m'], r['n'], r['k'], 'eq', r['bitwise_equal'], 'ref', r['ref_us'], 'ours', r['our_us'], 'e2e', r['ref_e2e_us'], r['our_e2e_us'])
"
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-1500
