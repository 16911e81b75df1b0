#!/bin/bash
# steady-state per-op timings + ncu full captures of the three dominant kernels on isolated ops
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 600 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops.txt 2>&1
timeout 600 python scripts/prof_ops.py --reps 40 --impl tc1 --only lin_k256_n256,lin_k256_n2048_geglu,conv_l2_256,lin_k640_n640 > gpurun_out/prof_ops_tc1.txt 2>&1
for c in lin_k256_n256 lin_k256_n2048_geglu conv_l2_256; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc2 -s 1 -c 1 -f -o gpurun_out/ncu_$c \
      python scripts/prof_ops.py --reps 2 --only $c > gpurun_out/ncu_$c.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -f -o gpurun_out/ncu_attn_1024 \
    python scripts/prof_ops.py --reps 2 --only attn_1024 > gpurun_out/ncu_attn_1024.log 2>&1
cat gpurun_out/prof_ops.txt gpurun_out/prof_ops_tc1.txt
