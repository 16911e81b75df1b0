#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 4000 --csv --log-file gpurun_out/launches_warm.csv \
    python bench.py --steps 1 --warmup 0 --ddim-steps 2 --no-cpu-baseline --no-kernel-pass --no-graph > gpurun_out/ncu_list.log 2>&1
tail -c 300 gpurun_out/ncu_list.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-kernel-pass > gpurun_out/bench.log 2>&1; tail -c 400 gpurun_out/bench.log
