#!/bin/bash
# One gpurun call: per-kernel parity (each family in its own process so a trapped kernel cannot
# poison the others), network parity, bench, and ncu captures.  Logs -> gpurun_out/.
# Usage: bash scripts/gpu_check.sh [quick|full|prof]
MODE=${1:-full}
set +e
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
: > gpurun_out/summary.txt
run() { name=$1; shift; timeout ${TMO:-900} "$@" > gpurun_out/$name.log 2>&1; echo "$name exit $?" >> gpurun_out/summary.txt; tail -3 gpurun_out/$name.log >> gpurun_out/summary.txt; }
if [ "$MODE" != "prof" ]; then
run ops_misc   python -m pytest tests/test_gpu_ops.py -q -k "not gemm and not attention" -p no:cacheprovider
run ops_simt   python -m pytest tests/test_gpu_ops.py -q -k "gemm and simt" -p no:cacheprovider
run ops_tc     python -m pytest tests/test_gpu_ops.py -q -k "gemm and tc and not tc1" -p no:cacheprovider
run ops_tc1    python -m pytest tests/test_gpu_ops.py -q -k "gemm and tc1" -p no:cacheprovider
run ops_attn   python -m pytest tests/test_gpu_ops.py -q -k "attention" -p no:cacheprovider
run nets_tiny  python -m pytest tests/test_gpu_nets.py -q -s -k "tiny" -p no:cacheprovider
run nets_full  python -m pytest tests/test_gpu_nets.py -q -s -k "full or end_to_end" -p no:cacheprovider
run smoke      python __graft_entry__.py smoke
fi
if [ "$MODE" = "full" ]; then
run bench      python bench.py --steps 1 --warmup 1 --dump-ops gpurun_out/ops.csv --torch-cuda-baseline ${BENCH_ARGS}
fi
if [ "$MODE" = "bench" ]; then
run bench      python bench.py --steps 1 --warmup 1 --dump-ops gpurun_out/ops.csv --no-cpu-baseline ${BENCH_ARGS}
fi
if [ "$MODE" = "full" ] || [ "$MODE" = "prof" ]; then
TMO=1200 run ncu_list ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 --ddim-steps 2 --no-cpu-baseline --no-kernel-pass --no-graph
TMO=1200 run ncu_gemm ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 300 -c 4 -f -o gpurun_out/prof_gemm \
    python bench.py --steps 1 --warmup 0 --ddim-steps 1 --no-cpu-baseline --no-kernel-pass --no-graph
TMO=900 run ncu_attn ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 20 -c 2 -f -o gpurun_out/prof_attn \
    python bench.py --steps 1 --warmup 0 --ddim-steps 1 --no-cpu-baseline --no-kernel-pass --no-graph
fi
cat gpurun_out/summary.txt
