#!/bin/bash
# usage: gpu_quick.sh "<pytest -k expr or empty>" "<profile_kernels --only list or empty>" [bench]
mkdir -p gpurun_out
if [ -n "$1" ]; then
  timeout 900 python -m pytest tests -q -m gpu -x -k "$1" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
  tail -6 gpurun_out/pytest_quick.log | cut -c1-300
fi
if [ -n "$2" ]; then
  timeout 600 python tools/profile_kernels.py --time --only "$2" --out gpurun_out/kt_quick.json > gpurun_out/kt_quick.log 2>&1
  cut -c1-200 gpurun_out/kt_quick.log
fi
if [ "$3" == "bench" ]; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; tail -1 gpurun_out/bench_quick.log | cut -c1-330
fi
