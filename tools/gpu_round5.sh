#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python tools/bringup.py mcast > gpurun_out/mcast.log 2>&1; echo "mcast rc=$?"
grep '"mcast"' gpurun_out/mcast.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    if 'error' in r: print(r); continue
    c = r['cfg']; print(r['m'], r['n'], r['k'], 'bm', c['block_m'], 'cl', c['cluster'], 'st', c['num_stages'], 'eq', r['equal'], r['us'], 'us', r['tflops'])
"
tail -3 gpurun_out/mcast.log | cut -c1-400
