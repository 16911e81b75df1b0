#!/bin/bash
# Round-end style validation + the captures that go to profiles/.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python scripts/prof_small.py 50 > gpurun_out/prof_small.json 2>&1; cat gpurun_out/prof_small.json
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops.txt 2>&1; cat gpurun_out/prof_ops.txt
for c in conv_l2_256 lin_k1024_n256; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 1 -c 1 -f -o gpurun_out/ncu_$c python scripts/prof_ops.py --reps 2 --only $c > gpurun_out/ncu_$c.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -f -o gpurun_out/ncu_attn_1024 python scripts/prof_ops.py --reps 2 --only attn_1024 > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"ddim_step|stft_mel|gn_stats_col|gn_apply_col|ln_kernel" -c 12 -f -o gpurun_out/ncu_small python scripts/prof_small.py 1 > gpurun_out/ncu_small.log 2>&1
timeout 1200 python bench.py --dump-ops gpurun_out/ops.csv > gpurun_out/bench.log 2>&1; tail -c 2500 gpurun_out/bench.log
