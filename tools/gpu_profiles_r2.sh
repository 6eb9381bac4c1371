#!/usr/bin/env bash
# Round-2 evidence run (1 GPU): smoke, the bench line, ncu launch list of the bench command, ncu --set full of the top kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2p
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/smi.txt
timeout 600 python tools/tune.py swap2 > $O/tune_swap2.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.log
timeout 900 python bench.py > $O/bench.log 2> $O/bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > $O/bench_ref_arm.log 2> $O/bench_ref_arm.err; echo "bench ref rc=$?"; tail -c 600 $O/bench_ref_arm.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > $O/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
NCU="ncu --set full --clock-control none --import-source on"
prof() {  # name kernel-regex skip -- command
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 $NCU -k regex:$rx -s $skip -c 1 -f -o $O/$name "$@" > $O/prof_$name.log 2>&1; echo "$name rc=$?"
}
prof ours_4096 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 4096 4096 7168
prof ref_4096 sm100_fp8 1 python tools/prof_r2.py dense ref 4096 4096 7168
prof ours_64 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 64 4096 7168
prof ref_64 sm100_fp8 1 python tools/prof_r2.py dense ref 64 4096 7168
prof ours_128 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 128 4096 7168
prof ours_192 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 192 4096 7168
prof ours_512 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 512 4096 7168
prof ours_k2048 fp8_gemm_kernel 1 python tools/prof_r2.py dense ours 4096 7168 2048
prof ref_k2048 sm100_fp8 1 python tools/prof_r2.py dense ref 4096 7168 2048
prof ours_contig fp8_gemm_kernel 1 python tools/prof_r2.py contiguous ours 128
prof ref_contig sm100_fp8 1 python tools/prof_r2.py contiguous ref 128
prof ours_masked fp8_gemm_kernel 1 python tools/prof_r2.py masked ours 64
prof ours_quant per_token_cast 1 python tools/prof_r2.py quant 4096 7168
prof ep_dispatch dispatch_fused 1 python tools/prof_r2.py ep
prof ep_combine combine_gather 1 python tools/prof_r2.py ep
for f in $O/*.ncu-rep; do ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null; done
ncu -i $O/ours_4096.ncu-rep --page source --csv > $O/ours_4096.source.csv 2>/dev/null
ncu -i $O/ours_64.ncu-rep --page source --csv > $O/ours_64.source.csv 2>/dev/null
ncu -i $O/ours_k2048.ncu-rep --page source --csv > $O/ours_k2048.source.csv 2>/dev/null
find $O -name "*.ncu-rep" ! -name "ours_4096.ncu-rep" -delete
bash tools/gpu_sanitize.sh
du -sh $O; ls $O | head -60
