#!/bin/bash
# One gpurun call at the end of a round: full GPU suite, smoke, the default bench line, the launch list of the same
# command, a graph-replay timeline, the step breakdown and an ncu --set full capture of the prologue backward kernels.
# usage: gpu_final.sh <tag>      (files land in gpurun_out/<tag>_*)
T=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
tail -2 gpurun_out/${T}_smoke.log | cut -c1-200
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -1 gpurun_out/${T}_bench.json | cut -c1-400
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${T}_launches_raw.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-eager > gpurun_out/${T}_bench_ncu.log 2>&1; echo "launch list rc=$?"
timeout 120 python tools/timeline.py --graph --dump --out gpurun_out/${T}_timeline_graph.json > gpurun_out/${T}_timeline.log 2>&1; head -3 gpurun_out/${T}_timeline.log | cut -c1-300
timeout 120 python tools/step_breakdown.py > gpurun_out/${T}_step_breakdown.txt 2>&1; tail -5 gpurun_out/${T}_step_breakdown.txt | cut -c1-200
timeout 200 ncu --set full --clock-control none --import-source on -f -o gpurun_out/${T}_prologue -k regex:'stream_prologue_bwd' \
  python tools/profile_kernels.py --only stream_prologue_bwd > gpurun_out/${T}_prologue_ncu.log 2>&1; echo "ncu prologue rc=$?"
