#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 900 python tools/big_shapes.py > $OUT/big_shapes.log 2>&1; echo "rc=$?"; grep '^{' $OUT/big_shapes.log; tail -3 $OUT/big_shapes.log | cut -c1-300
