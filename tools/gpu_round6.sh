#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/stamps.py > gpurun_out/stamps.log 2>&1; echo "stamps rc=$?"; cat gpurun_out/stamps.log | cut -c1-700
timeout 600 python tools/bringup.py ref > gpurun_out/ref.log 2>&1; echo "ref rc=$?"
grep ref_vs_ours gpurun_out/ref.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['m'], r['n'], r['k'], 'bm', r['cfg']['block_m'], 'S', r['cfg']['num_splits'], 'eq', r['bitwise_equal'], r['mismatches'], 'ref', r['ref_us'], 'ours', r['our_us'], 'e2e', r['ref_e2e_us'], r['our_e2e_us'])
"
