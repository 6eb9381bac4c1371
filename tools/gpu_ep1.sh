mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ep_gpu.py -x -q -m gpu 2>&1 | tail -3
for ov in 0 1; do
  DGB200_EP_OVERLAP=$ov timeout 900 python bench.py --gpus 1 --workload ep --steps 10 --warmup 3 > gpurun_out/b_$ov.log 2>&1; echo "overlap=$ov rc=$?"; tail -1 gpurun_out/b_$ov.log | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print({k: d[k] for k in ('n_gpus', 'value', 'ms_per_step', 'dispatch_ms', 'gemm_ms', 'overlap')})
except Exception as e:
    print('no json')" ; grep -i "error\|Traceback" -A3 gpurun_out/b_$ov.log | head -8
done
