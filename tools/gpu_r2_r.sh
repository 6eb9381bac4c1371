#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 900 python tools/tune.py swap_mid > $OUT/tune_r_swap_mid.log 2>&1
grep -v -i warn $OUT/tune_r_swap_mid.log | tail -70
