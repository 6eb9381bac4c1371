#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 300 python tools/host_overhead.py > $OUT/host_overhead.log 2>&1; tail -2 $OUT/host_overhead.log
