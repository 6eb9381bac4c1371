#!/usr/bin/env bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for cs in auto 0; do
  echo "== DGB200_CSPLIT=$cs"
  if [ "$cs" = "auto" ]; then unset DGB200_CSPLIT; else export DGB200_CSPLIT=$cs; fi
  timeout 600 python tools/bringup.py ref > /tmp/o.log 2>&1; grep ref_vs_ours /tmp/o.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print(r['m'], r['n'], r['k'], 'bm', r['cfg']['block_m'], 'cl', r['cfg']['cluster'], 'S', r['cfg']['num_splits'], 'cs', r['cfg'].get('cluster_split'), 'eq', r['bitwise_equal'], r['mismatches'], 'ref', r['ref_us'], 'ours', r['our_us'])
"; grep -i "error\|Traceback" /tmp/o.log | head -5
done
unset DGB200_CSPLIT
for cs in 4; do echo "CSPLIT=$cs"; DGB200_CSPLIT=$cs python tools/stamps.py 2>&1 | cut -c1-900;  DGB200_CSPLIT=$cs python tools/stamps.py --cold 2>&1 | cut -c1-900; done
