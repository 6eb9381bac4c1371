#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 600 python tools/tune.py tall > $OUT/tune_x_tall.log 2>&1; grep -v -i warn $OUT/tune_x_tall.log
