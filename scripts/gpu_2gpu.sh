#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --ddim-steps 20 --no-kernel-pass --no-cpu-baseline > gpurun_out/bench_2gpu.log 2>&1
tail -c 2500 gpurun_out/bench_2gpu.log
