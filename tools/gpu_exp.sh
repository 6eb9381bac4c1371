#!/usr/bin/env bash
for v in deepgemm_b200/lib/exp/libdgb200_v8.so deepgemm_b200/lib/exp/libdgb200_v15.so; do
  echo "== LIB=$v"
  DGB200_LIB=$v timeout 600 python tools/bringup.py ref > /tmp/o.log 2>&1; grep ref_vs_ours /tmp/o.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['m'], r['n'], r['k'], 'bm', r['cfg']['block_m'], 'ref', r['ref_us'], 'ours', r['our_us'])
"
done
