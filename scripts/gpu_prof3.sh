#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "(gemm and tc and not tc1) or attention" -p no:cacheprovider > gpurun_out/ops_tc.log 2>&1; tail -2 gpurun_out/ops_tc.log
timeout 600 python -m pytest tests/test_gpu_nets.py -q -s -k "tiny" -p no:cacheprovider > gpurun_out/nets_tiny.log 2>&1; tail -2 gpurun_out/nets_tiny.log
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops.txt 2>&1
cat gpurun_out/prof_ops.txt
timeout 300 python scripts/prof_ops.py --reps 3 --dbg 128 --only conv_l2_256,lin_k256_n2048_geglu > gpurun_out/timeline.txt 2>&1
head -30 gpurun_out/timeline.txt; grep -A5 "tile |" gpurun_out/timeline.txt
