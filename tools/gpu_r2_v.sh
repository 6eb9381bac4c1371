#!/usr/bin/env bash
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
OUT=gpurun_out/r2; mkdir -p $OUT
BENCH_FORCE_RETIME=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 5 --warmup 3 --no-extras --no-cpu-baseline > $OUT/bench_retime_$N.log 2> $OUT/bench_retime_$N.err
echo "rc=$?"; tail -1 $OUT/bench_retime_$N.log | cut -c1-900
BENCH_FORCE_RETIME=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus $N --impl reference --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_retime_ref_$N.log 2> $OUT/bench_retime_ref_$N.err
echo "rc=$?"; tail -1 $OUT/bench_retime_ref_$N.log | cut -c1-700
