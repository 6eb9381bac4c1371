#!/usr/bin/env bash
# what the driver runs at round end on one fresh B200: the GPU test tier, smoke(), both bench arms
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2d; mkdir -p $O
( time timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider ) > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -c 300 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
( time timeout 900 python bench.py --impl reference ) > $O/bench_ref.log 2> $O/bench_ref.err; echo "ref rc=$?"
( time timeout 900 python bench.py ) > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"; tail -4 $O/bench.err
python - <<'PY'
import json
for f in ('bench_ref.log', 'bench.log'):
    j = json.loads([l for l in open('gpurun_out/r2d/' + f) if l.startswith('{')][-1])
    print(f, j['value'], j['ms_per_step'], j['e2e']['value'], j.get('gpu_launches'), j['config'].get('retimed'))
PY
