#!/bin/bash
# GEMM parity (incl. the weight-stationary variant) + per-kernel timing table, with and without the WS variant.
mkdir -p gpurun_out
timeout 600 python tools/gpu_kernel_check.py --case gemm > gpurun_out/gemm_check.log 2>&1; echo "rc=$?" >> gpurun_out/gemm_check.log
tail -5 gpurun_out/gemm_check.log | cut -c1-600
timeout 600 python tools/profile_kernels.py --time --only gemm_ --out gpurun_out/kt_ws.json > gpurun_out/kt_ws.log 2>&1
TMP_B200_GEMM_NO_WS=1 timeout 600 python tools/profile_kernels.py --time --only gemm_ --out gpurun_out/kt_nows.json > gpurun_out/kt_nows.log 2>&1
paste -d'\n' gpurun_out/kt_ws.log gpurun_out/kt_nows.log | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gemm.log 2>&1; tail -1 gpurun_out/bench_gemm.log | cut -c1-300
