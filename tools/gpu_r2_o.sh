#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
( time timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_ep_gpu.py::test_multi_gpu_peer_dispatch_under_torchrun -p no:cacheprovider ) > $OUT/pytest_gpu_o.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu_o.log; tail -c 400 $OUT/pytest_gpu_o.log
( time timeout 900 python bench.py ) > $OUT/bench_o.log 2> $OUT/bench_o.err; echo "bench rc=$?"
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r2/bench_o.log').read().strip().splitlines()[-1])
print(j['value'], j['e2e']['value'], json.dumps(j['decode_chain'])[:700])
for r in j['vs_reference_kernel']['per_shape']: print(r['m'], r['ours_us'], r['ref_kernel_us'], r['ours_kineto_us'], r['ref_kineto_us'])
PY
