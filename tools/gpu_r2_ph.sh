#!/usr/bin/env bash
N=${1:-8}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 tools/ep_phases.py ) > $OUT/ep_phases_$N.log 2>&1
echo "rc=$?" >> $OUT/ep_phases_$N.log
grep '^{' $OUT/ep_phases_$N.log | head -3
