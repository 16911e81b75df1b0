#!/bin/bash
# Round-2 opener: validate the two switches that were written without GPU time at the end of round 1
# (ALDM_BN256=1: 128 x 256 GEMM tiles; ALDM_GN_FUSED=1: single-launch GroupNorm), then A/B them on the step loop.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
B="python bench.py --steps 2 --warmup 3 --ddim-steps 50 --no-cpu-baseline --no-kernel-pass"
pick() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        j=json.loads(l); print(sys.argv[1], "clips/s", round(j["value"],4), "ms/ddim", round(j["breakdown"]["ms_per_ddim_step"],3))
PY
}
for sw in ALDM_BN256 ALDM_GN_FUSED; do
  echo "== $sw=1"
  env $sw=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -p no:cacheprovider > gpurun_out/pytest_$sw.log 2>&1; tail -3 gpurun_out/pytest_$sw.log
  env $sw=1 timeout 300 $B > gpurun_out/ab_$sw.log 2>&1; pick gpurun_out/ab_$sw.log
done
timeout 300 $B > gpurun_out/ab_default.log 2>&1; pick gpurun_out/ab_default.log
# compiled-out experiment: A operand of the linear layers by tensor-map TMA (rebuilds the library on the box only)
echo "== ALDM_TMA_A=1 (build with -DALDM_EXPERIMENTAL_TMA)"
ALDM_BUILD_EXPERIMENTAL=1 python -c "from audioldm2_b200 import _lib; _lib.build(force=True)" > gpurun_out/build_exp.log 2>&1; tail -1 gpurun_out/build_exp.log
ALDM_TMA_A=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -p no:cacheprovider > gpurun_out/pytest_tma.log 2>&1; tail -3 gpurun_out/pytest_tma.log
ALDM_TMA_A=1 timeout 300 $B > gpurun_out/ab_tma.log 2>&1; pick gpurun_out/ab_tma.log
ALDM_TMA_A=1 timeout 300 python scripts/prof_ops.py --reps 40 2>&1 | tee gpurun_out/prof_ops_tma.txt
python -c "from audioldm2_b200 import _lib; _lib.build(force=True)" > /dev/null 2>&1
ALDM_BN256=1 timeout 300 python scripts/prof_ops.py --reps 40 2>&1 | tee gpurun_out/prof_ops_bn256.txt
