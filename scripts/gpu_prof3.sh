#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 300 python scripts/prof_ops.py --reps 3 --dbg 128 --only conv_l2_256,lin_k256_n2048_geglu > gpurun_out/timeline.txt 2>&1
timeout 300 python scripts/prof_ops.py --reps 3 --dbg 143 --only conv_l2_256 >> gpurun_out/timeline.txt 2>&1
cat gpurun_out/timeline.txt
