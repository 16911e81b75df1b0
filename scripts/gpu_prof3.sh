#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "(gemm and tc and not tc1) or attention or groupnorm" -p no:cacheprovider > gpurun_out/ops_tc.log 2>&1; tail -2 gpurun_out/ops_tc.log
timeout 600 python -m pytest tests/test_gpu_nets.py -q -s -k "tiny or full_vs" -p no:cacheprovider > gpurun_out/nets_tiny.log 2>&1; tail -2 gpurun_out/nets_tiny.log
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops.txt 2>&1
cat gpurun_out/prof_ops.txt
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --dump-ops gpurun_out/ops.csv > gpurun_out/bench.log 2>&1; tail -c 700 gpurun_out/bench.log
