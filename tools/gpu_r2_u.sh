#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
FUZZ_SECONDS=300 timeout 900 python tools/parity_fuzz.py > $OUT/parity_fuzz.log 2>&1; echo "fuzz rc=$?"; tail -3 $OUT/parity_fuzz.log | cut -c1-1500
