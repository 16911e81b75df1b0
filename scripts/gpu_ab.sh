#!/bin/bash
# A/B of launch/pipeline switches on the real step loop (env-controlled), after a parity pass.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
ALDM_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -p no:cacheprovider > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log; grep -m2 "aldm\]" gpurun_out/pytest_ab.log
B="python bench.py --steps 2 --warmup 3 --ddim-steps 50 --no-cpu-baseline --no-kernel-pass"
pick() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        j=json.loads(l); print(sys.argv[1], "clips/s", round(j["value"],4), "ms/ddim", round(j["breakdown"]["ms_per_ddim_step"],3))
PY
}
ALDM_PDL=0 ALDM_ATTN_STAGES=2 timeout 300 $B > gpurun_out/ab_base.log 2>&1; pick gpurun_out/ab_base.log
ALDM_VERBOSE=1 ALDM_PDL=0 timeout 300 $B > gpurun_out/ab_attn3.log 2>&1; pick gpurun_out/ab_attn3.log; grep -m1 "aldm\]" gpurun_out/ab_attn3.log
ALDM_PDL=0 timeout 300 $B --no-graph > gpurun_out/ab_eager.log 2>&1; pick gpurun_out/ab_eager.log
timeout 300 $B --no-graph > gpurun_out/ab_eager_pdl.log 2>&1; pick gpurun_out/ab_eager_pdl.log
timeout 300 $B > gpurun_out/ab_all.log 2>&1; pick gpurun_out/ab_all.log
timeout 300 python scripts/prof_ops.py --reps 40 > gpurun_out/prof_ops_ab.txt 2>&1; cat gpurun_out/prof_ops_ab.txt
