#!/usr/bin/env bash
# Round-2 multi-GPU pass: usage tools/gpu_r2_mg.sh N   (run under gpurun --gpus N)
set -u
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2
mkdir -p $OUT
nvidia-smi -L > $OUT/mg_${N}_smi.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/ep_check.py ) > $OUT/ep_check_$N.log 2>&1
echo "ep_check rc=$?" >> $OUT/ep_check_$N.log
tail -n 12 $OUT/ep_check_$N.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 ) > $OUT/bench_mg_$N.log 2> $OUT/bench_mg_$N.err
echo "bench rc=$?" >> $OUT/bench_mg_$N.err
tail -c 600 $OUT/bench_mg_$N.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --impl reference --steps 10 --warmup 3 --no-cpu-baseline ) > $OUT/bench_mg_ref_$N.log 2> $OUT/bench_mg_ref_$N.err
echo "ref arm rc=$?" >> $OUT/bench_mg_ref_$N.err
python - <<'PY'
import json, sys, glob
for f in sorted(glob.glob('gpurun_out/r2/bench_mg_*.log')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('impl'), d.get('n_gpus'), d.get('value'), json.dumps(d.get('ep'))[:900])
    except Exception as e:
        print(f, 'unparsable', e)
PY
