#!/usr/bin/env bash
# Round-1 evidence run (1 GPU): tests, smoke, bench, ncu launch list of the bench command, ncu --set full of the top kernels.
O=gpurun_out/r1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.log 2>&1; echo "bench rc=$?"; tail -1 $O/bench.log | cut -c1-1500
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_arm.log 2>&1; echo "bench ref rc=$?"; tail -1 $O/bench_ref_arm.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 900 python bench.py --workload contiguous --steps 5 --warmup 3 > $O/bench_contig.log 2>&1; echo "contig rc=$?"; tail -1 $O/bench_contig.log | cut -c1-900
timeout 900 python bench.py --workload masked --steps 5 --warmup 3 > $O/bench_masked.log 2>&1; echo "masked rc=$?"; tail -1 $O/bench_masked.log | cut -c1-900
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $O/ours_4096 python tools/prof_one.py ours 4096 4096 7168 > $O/prof_ours_4096.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:sm100_fp8 -s 1 -c 1 -f -o $O/ref_4096 python tools/prof_one.py ref 4096 4096 7168 > $O/prof_ref_4096.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $O/ours_64 python tools/prof_one.py ours 64 4096 7168 > $O/prof_ours_64.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:sm100_fp8 -s 1 -c 1 -f -o $O/ref_64 python tools/prof_one.py ref 64 4096 7168 > $O/prof_ref_64.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $O/ours_512 python tools/prof_one.py ours 512 4096 7168 > $O/prof_ours_512.log 2>&1; echo "rc=$?"
timeout 600 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $O/ours_128 python tools/prof_one.py ours 128 4096 7168 > $O/prof_ours_128.log 2>&1; echo "rc=$?"
timeout 900 $NCU -k regex:fp8_gemm_kernel -s 1 -c 1 -f -o $O/ours_contig python tools/prof_grouped.py ours 48 256 > $O/prof_ours_contig.log 2>&1; echo "rc=$?"
timeout 900 $NCU -k regex:"scatter_kernel|bucket_kernel|exchange_kernel|wait_kernel" -s 8 -c 4 -f -o $O/ep_dispatch python bench.py --workload ep --steps 2 --warmup 3 > $O/prof_ep.log 2>&1; echo "rc=$?"
for f in $O/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
ncu -i $O/ours_4096.ncu-rep --page source --csv > $O/ours_4096.source.csv 2>/dev/null
ncu -i $O/ours_64.ncu-rep --page source --csv > $O/ours_64.source.csv 2>/dev/null
find $O -name "*.ncu-rep" ! -name "ours_4096.ncu-rep" -delete
du -sh $O; ls -la $O
