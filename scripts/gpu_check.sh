#!/bin/bash
# One gpurun call: per-kernel parity (each family in its own process so a trapped kernel cannot
# poison the others), network parity, a short bench and an ncu launch list.  Logs -> gpurun_out/.
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log >> gpurun_out/summary.txt; }
: > gpurun_out/summary.txt
run ops_misc   python -m pytest tests/test_gpu_ops.py -q -k "not gemm" -p no:cacheprovider
run ops_simt   python -m pytest tests/test_gpu_ops.py -q -k "gemm and simt" -p no:cacheprovider
run ops_tc     python -m pytest tests/test_gpu_ops.py -q -k "gemm and tc" -p no:cacheprovider
run nets_tiny  python -m pytest tests/test_gpu_nets.py -q -s -k "tiny" -p no:cacheprovider
run nets_full  python -m pytest tests/test_gpu_nets.py -q -s -k "full or end_to_end" -p no:cacheprovider
run smoke      python __graft_entry__.py smoke
run bench      python bench.py --steps 1 --warmup 1 ${BENCH_ARGS}
cat gpurun_out/summary.txt
