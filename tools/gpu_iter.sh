#!/bin/bash
# One gpurun call of an optimisation iteration: full GPU test suite, kernel timing table, bench (1 GPU).
# usage: gpu_iter.sh [profile_kernels --only list]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python tools/profile_kernels.py --time --only "$1" --out gpurun_out/kernel_times.json > gpurun_out/kernel_times.log 2>&1
cut -c1-200 gpurun_out/kernel_times.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
