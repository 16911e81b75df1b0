#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ops.py -q -k "gemm and tc and not tc1" -p no:cacheprovider > gpurun_out/ops_tc.log 2>&1; tail -2 gpurun_out/ops_tc.log
CASES=lin_k256_n256,lin_k256_n768_qkv,lin_k1024_n256,lin_k256_n2048_geglu,conv_l2_256,lin_k640_n640
for dbg in 0 1 2 3 4 8 12 15; do
  echo "== dbg $dbg" >> gpurun_out/prof_dbg.txt
  timeout 300 python scripts/prof_ops.py --reps 40 --only $CASES --dbg $dbg >> gpurun_out/prof_dbg.txt 2>&1
done
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops.txt 2>&1
cat gpurun_out/prof_dbg.txt gpurun_out/prof_ops.txt
