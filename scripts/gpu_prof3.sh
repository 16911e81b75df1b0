#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/prof_dbg.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "(gemm and tc and not tc1) or attention" -p no:cacheprovider > gpurun_out/ops_tc.log 2>&1; tail -2 gpurun_out/ops_tc.log
timeout 600 python -m pytest tests/test_gpu_nets.py -q -s -k "tiny" -p no:cacheprovider > gpurun_out/nets_tiny.log 2>&1; tail -2 gpurun_out/nets_tiny.log
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops.txt 2>&1; timeout 300 python scripts/prof_ops.py --reps 40 --dbg 32 --only conv_l2_256,lin_k1024_n256 >> gpurun_out/prof_ops.txt 2>&1
cat gpurun_out/prof_ops.txt
echo
