#!/bin/bash
# What the driver runs at round end, in one call: single-process GPU tests, smoke, default bench (+ per-op table), micro-benchmarks.
set +e
cd "$(dirname "$0")/.."
T=${1:-final}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$T.log 2>&1; tail -1 gpurun_out/smoke_$T.log
timeout 1200 python bench.py --dump-ops gpurun_out/ops_$T.csv > gpurun_out/bench_$T.log 2>&1; tail -c 3000 gpurun_out/bench_$T.log
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops_$T.txt 2>&1; cat gpurun_out/prof_ops_$T.txt
