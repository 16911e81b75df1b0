#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ALDM_BN256=1 timeout 25 python -m pytest tests/test_gpu_ops.py -x -q -k "gemm and tc and not tc1 and (bn256 or geglu or dual_out)" -p no:cacheprovider > gpurun_out/t_bn256.log 2>&1; tail -4 gpurun_out/t_bn256.log
ALDM_GN_FUSED=1 timeout 20 python -m pytest tests/test_gpu_ops.py -x -q -k "groupnorm or prep" -p no:cacheprovider > gpurun_out/t_gn.log 2>&1; tail -4 gpurun_out/t_gn.log
