#!/usr/bin/env bash
# last validation of the final build on one GPU: full suite, sanitizers over the extended workload, smoke, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2f; mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -c 500 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
for tool in memcheck synccheck; do
  ( time timeout 900 env LONG_K=1 compute-sanitizer --tool $tool --target-processes=application-only --report-api-errors no --error-exitcode 7 \
      python tools/sanitize_workload.py ) > $O/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/sanitize_$tool.log
  grep -E "ERROR SUMMARY|sanitize workload done" $O/sanitize_$tool.log | tail -2
done
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"
