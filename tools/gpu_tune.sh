mkdir -p gpurun_out
timeout 1200 python tools/tune.py big > gpurun_out/tune_big.log 2>&1; tail -60 gpurun_out/tune_big.log
timeout 1200 python tools/tune.py small > gpurun_out/tune_small.log 2>&1; tail -130 gpurun_out/tune_small.log
