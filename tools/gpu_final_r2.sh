#!/usr/bin/env bash
# Round-2 final evidence run (1 GPU): full GPU suite, smoke, both bench arms, ncu launch list of the bench command, BF16 /
# k-grouped A/B, sanitizers over the extended workload.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2f
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt
( time timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -c 600 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > $O/bench_ref_arm.log 2> $O/bench_ref_arm.err; echo "bench ref rc=$?"; tail -c 400 $O/bench_ref_arm.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 300 python tools/bf16_bench.py > $O/bf16_bench.log 2>&1
KG_SWEEP=128,192 timeout 400 python tools/kgrouped_bench.py > $O/kgrouped_bench.log 2>&1
timeout 400 python tools/tune.py ab_small > $O/tune_ab_small.log 2>&1
OUTSAVE=$O
for tool in memcheck synccheck; do
  ( time timeout 900 env LONG_K=1 compute-sanitizer --tool $tool --target-processes=application-only --report-api-errors no --error-exitcode 7 \
      python tools/sanitize_workload.py ) > $O/sanitize_$tool.log 2>&1
  echo "$tool rc=$?" | tee -a $O/sanitize_$tool.log
  grep -E "ERROR SUMMARY|sanitize workload done" $O/sanitize_$tool.log | tail -3
done
du -sh $O; ls $O
