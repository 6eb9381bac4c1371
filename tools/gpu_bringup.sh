#!/usr/bin/env bash
# One gpurun call: bring-up correctness, reference A/B, sweep. Each stage has its own timeout so a hang costs one stage.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python tools/bringup.py dense --quick > gpurun_out/dense_quick.log 2>&1; rc=$?; echo "dense_quick rc=$rc"
tail -8 gpurun_out/dense_quick.log
if [ $rc -ne 0 ]; then echo "bring-up failed; stopping"; exit 1; fi
timeout 600 python tools/bringup.py dense > gpurun_out/dense.log 2>&1; echo "dense rc=$?"
tail -3 gpurun_out/dense.log
timeout 900 python tools/bringup.py ref > gpurun_out/ref.log 2>&1; echo "ref rc=$?"
tail -12 gpurun_out/ref.log
timeout 600 python tools/bringup.py sweep > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
grep -c sweep gpurun_out/sweep.log
