#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py -x -q -p no:cacheprovider > gpurun_out/pytest_ab.log 2>&1; tail -3 gpurun_out/pytest_ab.log
B="python bench.py --steps 2 --warmup 3 --ddim-steps 50 --no-cpu-baseline --no-kernel-pass"
pick() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        j=json.loads(l); print(sys.argv[1], "clips/s", round(j["value"],4), "ms/ddim", round(j["breakdown"]["ms_per_ddim_step"],3))
PY
}
timeout 300 $B > gpurun_out/ab_all.log 2>&1; pick gpurun_out/ab_all.log
timeout 300 python scripts/prof_ops.py --reps 40 --only lin_k256_n2048_geglu,lin_k640_n5120_geglu,gn_silu_l1,gn_silu_l4 2>&1 | tail -4
