#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/bringup.py ref > /tmp/o.log 2>&1; grep ref_vs_ours /tmp/o.log | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l)
    print(r['m'], r['n'], r['k'], 'bm', r['cfg']['block_m'], 'cl', r['cfg']['cluster'], 'S', r['cfg']['num_splits'], 'cs', r['cfg'].get('cluster_split'), 'eq', r['bitwise_equal'], r['mismatches'], 'ref', r['ref_us'], 'ours', r['our_us'])
"; grep -i "error\|Traceback" /tmp/o.log | head -5
timeout 1200 python tools/bringup.py grouped > gpurun_out/grouped.log 2>&1; grep "contiguous\|masked" gpurun_out/grouped.log | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l)
    except Exception: continue
    print(r['test'], r.get('mean_m'), 'bm', r['cfg']['block_m'], 'eq', r.get('bitwise_equal', r.get('valid_rows_bitwise_equal')), 'ref', r['ref_us'], 'ours', r['our_us'], 'graph', r.get('our_graph_us'))
"
