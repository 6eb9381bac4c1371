#!/usr/bin/env bash
mkdir -p gpurun_out
for e in 0 12; do
  echo "== DGB200_EARLY=$e"
  DGB200_EARLY=$e timeout 600 python tools/bringup.py ref 2>&1 | grep ref_vs_ours | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['m'], r['n'], r['k'], 'bm', r['cfg']['block_m'], 'ref', r['ref_us'], 'ours', r['our_us'])
"
done
for e in 0 12; do
DGB200_EARLY=$e timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct -k regex:fp8_gemm_kernel -s 1 -c 2 python tools/prof_one.py ours 64 4096 7168 2>&1 | grep -E "duration|dram__bytes|hit_rate"
done
timeout 300 python tools/stamps.py 2>&1 | grep '"m": 64' | cut -c1-700
