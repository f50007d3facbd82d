#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_swin_feed_gpu.py -q -m gpu -x > gpurun_out/pytest_swin.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_swin.log
tail -25 gpurun_out/pytest_swin.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_swin.log 2>&1; tail -2 gpurun_out/bench_swin.log | cut -c1-400
