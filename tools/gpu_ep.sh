#!/usr/bin/env bash
# usage: gpu_ep.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 python -m pytest tests/test_ep_gpu.py -x -q -m gpu > gpurun_out/test_ep_$N.log 2>&1; echo "pytest ep rc=$?"; tail -15 gpurun_out/test_ep_$N.log
timeout 600 $RUN tools/ep_check.py > gpurun_out/ep_check_$N.log 2>&1; echo "ep_check rc=$?"; grep "ep check" gpurun_out/ep_check_$N.log | head -8; grep -i "error\|assert" gpurun_out/ep_check_$N.log | head -10
timeout 900 $RUN bench.py --gpus $N --workload ep --steps 5 --warmup 3 > gpurun_out/bench_ep_$N.log 2>&1; echo "bench ep rc=$?"; tail -1 gpurun_out/bench_ep_$N.log | cut -c1-2200
if [ "$N" = "2" ]; then
  timeout 900 python bench.py --gpus 1 --workload ep --steps 5 --warmup 3 > gpurun_out/bench_ep_1.log 2>&1; echo "bench ep1 rc=$?"; tail -1 gpurun_out/bench_ep_1.log | cut -c1-2200
fi
