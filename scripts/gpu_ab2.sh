#!/bin/bash
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python __graft_entry__.py > gpurun_out/build.log 2>&1
ALDM_VERBOSE=1 timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "attention" -p no:cacheprovider > gpurun_out/pytest_ab.log 2>&1; tail -2 gpurun_out/pytest_ab.log; grep -m2 "aldm\]" gpurun_out/pytest_ab.log
timeout 600 python -m pytest tests/test_gpu_nets.py -x -q -p no:cacheprovider > gpurun_out/pytest_ab2.log 2>&1; tail -2 gpurun_out/pytest_ab2.log
B="python bench.py --steps 2 --warmup 3 --ddim-steps 50 --no-cpu-baseline --no-kernel-pass"
pick() { python - "$1" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{"metric"'):
        j=json.loads(l); print(sys.argv[1], "clips/s", round(j["value"],4), "ms/ddim", round(j["breakdown"]["ms_per_ddim_step"],3))
PY
}
ALDM_ATTN_STAGES=2 timeout 300 $B > gpurun_out/ab_attn2.log 2>&1; pick gpurun_out/ab_attn2.log
timeout 300 $B > gpurun_out/ab_all.log 2>&1; pick gpurun_out/ab_all.log
timeout 300 python scripts/prof_ops.py --reps 40 --only attn_1024,attn_256,attn_64 2>&1 | tee gpurun_out/prof_attn3.txt
for d in 1 2 4 8 3 7; do echo "dbg $d"; timeout 120 python scripts/prof_ops.py --reps 40 --dbg $d --only lin_k256_n256,lin_k640_n640,lin_k1024_n256,lin_k256_n2048_geglu 2>&1 | tail -4; done | tee gpurun_out/prof_dbg.txt
timeout 120 python scripts/prof_ops.py --reps 10 --dbg 128 --only lin_k256_n256,lin_k640_n640 > gpurun_out/timeline_small.txt 2>&1
