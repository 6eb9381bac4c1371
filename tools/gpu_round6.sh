#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for b in 1 0; do
echo "== DGB200_BALANCE=$b"
DGB200_BALANCE=$b timeout 600 python tools/bringup.py ref > gpurun_out/ref.log 2>&1; echo "ref rc=$?"
grep ref_vs_ours gpurun_out/ref.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if r['m'] < 512: continue
    print(r['m'], r['n'], r['k'], 'bm', r['cfg']['block_m'], 'st', r['cfg']['num_stages'], 'eq', r['bitwise_equal'], r['mismatches'], 'ref', r['ref_us'], 'ours', r['our_us'])
"
done
