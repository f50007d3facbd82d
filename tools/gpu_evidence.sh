#!/bin/bash
# One gpurun call collecting the round's evidence: ncu --set full of the hot kernels (with SASS/source), the attention
# sweep of BASELINE config 5, the step breakdown, the bench line and the launch list of the same command.
# usage: gpu_evidence.sh [profile_kernels --only list]
mkdir -p gpurun_out
ONLY=${1:-attn_fwd_T1005,attn_bwd_T1005,gemm_qkv_T1005,gemm_ffn1_T1005,gemm_ffn2_T1005,wgradb_ffn1_T1005,umse}
timeout 600 ncu --set full --clock-control none --import-source on -f -o gpurun_out/prof \
  -k regex:'attn_fwd_kernel|attn_bwd_kernel|gemm_tn|gemm_wgrad|umse' \
  python tools/profile_kernels.py --only "$ONLY" > gpurun_out/prof.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/profile_kernels.py --time --sweep --out gpurun_out/attn_sweep.json > gpurun_out/attn_sweep.log 2>&1
cut -c1-160 gpurun_out/attn_sweep.log | tail -30
timeout 300 python tools/profile_kernels.py --time --out gpurun_out/kernel_times.json > gpurun_out/kernel_times.log 2>&1
timeout 300 python tools/step_breakdown.py > gpurun_out/step_breakdown.log 2>&1; tail -25 gpurun_out/step_breakdown.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-1500
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "launch list rc=$?"
