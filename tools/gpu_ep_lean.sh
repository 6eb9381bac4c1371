#!/usr/bin/env bash
# usage: gpu_ep_lean.sh N  (ep_check + bench ep, both modes; no pytest)
N=${1:-4}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 300 $RUN tools/ep_check.py > gpurun_out/ep_check_$N.log 2>&1; echo "ep_check rc=$?"; grep -c "ep check ok" gpurun_out/ep_check_$N.log; grep -i "error\|assert" gpurun_out/ep_check_$N.log | head -5
for ov in 0 1; do
DGB200_EP_OVERLAP=$ov timeout 300 $RUN bench.py --gpus $N --workload ep --steps 10 --warmup 3 > gpurun_out/bench_ep_${N}_ov$ov.log 2>&1; echo "bench ep overlap=$ov rc=$?"; tail -1 gpurun_out/bench_ep_${N}_ov$ov.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('n_gpus', 'value', 'ms_per_step', 'dispatch_ms', 'gemm_ms', 'dispatch_alltoall_baseline_ms', 'overlap', 'tflops')})"
done
