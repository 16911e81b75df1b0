#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/prof_dbg.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
CASES=lin_k256_n2048_geglu,lin_k1024_n256,conv_l2_256,lin_k256_n768_qkv
for dbg in 15 47 79 111; do
  echo "== dbg $dbg" >> gpurun_out/prof_dbg.txt
  timeout 300 python scripts/prof_ops.py --reps 40 --only $CASES --dbg $dbg >> gpurun_out/prof_dbg.txt 2>&1
done
cat gpurun_out/prof_dbg.txt
