#!/usr/bin/env bash
# compute-sanitizer memcheck + synccheck + racecheck over tools/sanitize_workload.py (cf. reference tests/test_sanitizer.py:52-79)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
for tool in memcheck synccheck racecheck; do
  extra=LONG_K=1; [ $tool = racecheck ] && extra=LONG_K_LAST=1
  ( time timeout 900 env $extra compute-sanitizer --tool $tool --target-processes=application-only --report-api-errors no --error-exitcode 7 \
      python tools/sanitize_workload.py ) > $OUT/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $OUT/sanitize_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload done" $OUT/sanitize_$tool.log | tail -3
done
