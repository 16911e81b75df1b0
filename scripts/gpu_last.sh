#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 200 python -m pytest tests/test_gpu_zz_engine_abi.py -x -q -p no:cacheprovider > gpurun_out/pytest_engine.log 2>&1; tail -15 gpurun_out/pytest_engine.log
timeout 200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_last.log 2>&1; tail -3 gpurun_out/pytest_gpu_last.log
