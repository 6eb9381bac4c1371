#!/usr/bin/env bash
# usage: gpu_ep.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 600 python -m pytest tests/test_ep_gpu.py -x -q -m gpu > gpurun_out/test_ep_$N.log 2>&1; echo "pytest ep rc=$?"; tail -15 gpurun_out/test_ep_$N.log
timeout 600 $RUN tools/ep_check.py > gpurun_out/ep_check_$N.log 2>&1; echo "ep_check rc=$?"; grep "ep check" gpurun_out/ep_check_$N.log | head -8; grep -i "error\|assert" gpurun_out/ep_check_$N.log | head -10
for ov in 1 0; do
DGB200_EP_OVERLAP=$ov timeout 900 $RUN bench.py --gpus $N --workload ep --steps 10 --warmup 3 > gpurun_out/bench_ep_${N}_ov$ov.log 2>&1; echo "bench ep overlap=$ov rc=$?"; tail -1 gpurun_out/bench_ep_${N}_ov$ov.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('n_gpus', 'value', 'ms_per_step', 'dispatch_ms', 'gemm_ms', 'dispatch_alltoall_baseline_ms', 'overlap', 'tflops')})"
done
if [ "$N" = "2" ]; then
for ov in 1 0; do
  DGB200_EP_OVERLAP=$ov timeout 900 python bench.py --gpus 1 --workload ep --steps 10 --warmup 3 > gpurun_out/bench_ep_1_ov$ov.log 2>&1; echo "bench ep1 overlap=$ov rc=$?"; tail -1 gpurun_out/bench_ep_1_ov$ov.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('n_gpus', 'value', 'ms_per_step', 'dispatch_ms', 'gemm_ms', 'dispatch_alltoall_baseline_ms', 'overlap', 'tflops')})"
done
fi
