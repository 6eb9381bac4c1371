#!/usr/bin/env bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ep_gpu.py -x -q -m gpu > gpurun_out/test_ep_$N.log 2>&1; echo "pytest ep rc=$?"; tail -12 gpurun_out/test_ep_$N.log | cut -c1-300
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $RUN tools/ep_check.py > gpurun_out/ep_check_$N.log 2>&1; echo "ep_check rc=$?"; grep "ep check" gpurun_out/ep_check_$N.log | head -8; grep -i "error\|assert" gpurun_out/ep_check_$N.log | head -10
