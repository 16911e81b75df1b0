#!/bin/bash
# Round-end style validation + the captures that go to profiles/.  Usage: bash scripts/gpu_final.sh [tag]
set +e
cd "$(dirname "$0")/.."
T=${1:-final}
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$T.log 2>&1; tail -3 gpurun_out/pytest_gpu_$T.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$T.log 2>&1; tail -1 gpurun_out/smoke_$T.log
timeout 1200 python bench.py --dump-ops gpurun_out/ops_$T.csv > gpurun_out/bench_$T.log 2>&1; tail -c 3000 gpurun_out/bench_$T.log
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops_$T.txt 2>&1; cat gpurun_out/prof_ops_$T.txt
timeout 300 python scripts/prof_small.py 50 > gpurun_out/prof_small_$T.json 2>&1
# warm launch list of two eager DDIM steps (kernel shares) and DRAM traffic of every GEMM launch of one UNet evaluation
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$T.csv python bench.py --steps 1 --warmup 0 --ddim-steps 2 --no-cpu-baseline --no-kernel-pass --no-graph > gpurun_out/ncu_list_$T.log 2>&1
for c in conv_l2_256 lin_k256_n256 lin_k256_n2048_geglu; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 1 -c 1 -f -o gpurun_out/ncu_${T}_$c python scripts/prof_ops.py --reps 2 --only $c > gpurun_out/ncu_$c.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 1 -c 1 -f -o gpurun_out/ncu_${T}_attn_1024 python scripts/prof_ops.py --reps 2 --only attn_1024 > gpurun_out/ncu_attn.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"stft_mel|gn_stats_col|gn_apply_col|ln_kernel" -c 8 -f -o gpurun_out/ncu_${T}_small python scripts/prof_small.py 1 > gpurun_out/ncu_small.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$T.log 2>&1; tail -c 1200 gpurun_out/bench_ref_$T.log
