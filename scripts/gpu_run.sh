#!/bin/bash
# One parameterised GPU job runner (replaces the round-1 one-offs).  Usage (under gpurun, from the repo root):
#   bash scripts/gpu_run.sh tests            # pytest -m gpu
#   bash scripts/gpu_run.sh optests          # tests/test_gpu_ops.py only
#   bash scripts/gpu_run.sh step             # ms / DDIM step of the default configuration (scripts/step_time.py)
#   bash scripts/gpu_run.sh sweep            # ms / DDIM step for "lanes:ENV=.. ENV=..;..." configurations in $SWEEP
#   bash scripts/gpu_run.sh bench            # python bench.py $BENCH_ARGS -> gpurun_out/bench_<tag>.json
#   bash scripts/gpu_run.sh configs          # BASELINE configs C3 / C4 / C5, one timed batch each
#   bash scripts/gpu_run.sh ops              # steady-state micro-benchmarks of the dominant op shapes (scripts/prof_ops.py)
#   bash scripts/gpu_run.sh diag             # the GEMM with its roles switched off ($DIAG_BITS, $DIAG_CASES)
#   bash scripts/gpu_run.sh micro            # tcgen05.mma issue rates + per-role timelines of CTA 0 ($TL_CASES)
#   bash scripts/gpu_run.sh ncu-full         # one ncu --set full capture per case in $NCU_CASES -> raw-metric CSVs
#   bash scripts/gpu_run.sh ncu-list         # per-launch time + DRAM bytes of two DDIM steps (graph kernel nodes)
# Several jobs: bash scripts/gpu_run.sh tests step bench.  gpurun_out/ is capped at 64 MiB on the way back: no .ncu-rep files in it.
set +e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
python __graft_entry__.py > gpurun_out/build_$TAG.log 2>&1 || { tail -20 gpurun_out/build_$TAG.log; exit 1; }
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader > gpurun_out/smi_$TAG.txt
nproc > gpurun_out/nproc.txt
while [ $# -gt 0 ]; do
  job=$1; shift
  case $job in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1
      echo "== tests rc=$?"; tail -5 gpurun_out/pytest_$TAG.log ;;
    sweep)
      : > gpurun_out/sweep_$TAG.jsonl
      IFS=';' read -ra CFGS <<< "${SWEEP:-1:ALDM_TOKEN_PLANES=1;1:ALDM_TOKEN_PLANES=2}"      # "lanes:ENV=.. ENV=..;..."
      for cfg in "${CFGS[@]}"; do
        lanes=${cfg%%:*}; sw=${cfg#*:}
        env $sw timeout 300 python scripts/step_time.py --lanes $lanes --tag "lanes$lanes $sw" 2>gpurun_out/sweep_err.log | grep '^{' >> gpurun_out/sweep_$TAG.jsonl || tail -3 gpurun_out/sweep_err.log
      done
      echo "== sweep"; cat gpurun_out/sweep_$TAG.jsonl ;;
    bench)
      timeout 1500 python bench.py $BENCH_ARGS > gpurun_out/bench_$TAG.log 2>&1
      echo "== bench rc=$?"; grep '^{"metric"' gpurun_out/bench_$TAG.log > gpurun_out/bench_$TAG.json; tail -c 1500 gpurun_out/bench_$TAG.log ;;
    configs)      # BASELINE configs C3 / C4 (one GPU's share) / C5, short runs
      for c in "c3:--model audioldm_48k" "c4:--model audioldm2-full-large-1150k --no-torch-cuda-baseline" \
               "c5:--model audioldm_48k --mode sr_inpainting --batch 2"; do
        name=${c%%:*}; args=${c#*:}
        timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline $args > gpurun_out/bench_${name}_$TAG.log 2>&1
        echo "== bench $name rc=$?"; grep '^{"metric"' gpurun_out/bench_${name}_$TAG.log > gpurun_out/bench_${name}_$TAG.json; tail -c 600 gpurun_out/bench_${name}_$TAG.log
      done ;;
    refarm)
      timeout 1500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1
      echo "== reference arm rc=$?"; tail -c 1200 gpurun_out/bench_ref_$TAG.log ;;
    ops)
      for pl in 1 2; do timeout 600 python scripts/prof_ops.py --reps 40 --planes $pl > gpurun_out/prof_ops_${TAG}_p$pl.txt 2>&1; echo "== prof_ops planes=$pl"; cat gpurun_out/prof_ops_${TAG}_p$pl.txt; done ;;
    optests)
      timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_ops_$TAG.log 2>&1
      echo "== op tests rc=$?"; tail -15 gpurun_out/pytest_ops_$TAG.log ;;
    nettests)
      timeout 1500 python -m pytest tests/test_gpu_nets.py tests/test_gpu_zz_engine_abi.py -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_nets_$TAG.log 2>&1
      echo "== net tests rc=$?"; grep "rel L2\|passed\|failed\|FAILED\|Error" gpurun_out/pytest_nets_$TAG.log | tail -60 ;;
    ncu-full)     # one --set full capture per dominant op shape (scripts/prof_ops.py cases) + the small HBM kernels; only the
                  # raw-metric CSV comes back (gpurun_out is capped at 64 MiB: the .ncu-rep files stay on the box)
      for c in ${NCU_CASES:-lin_k1024_n256 lin_k256_n2048_geglu lin_k640_n640 conv_l2_256 attn_1024 ln_16384x256 gn_silu_l1}; do
        case $c in attn*) pat='attention_tc';; ln*) pat='ln_kernel';; gn*) pat='gn_fused|gn_apply_col';; *) pat='gemm_tc3';; esac
        timeout 600 ncu --set full --clock-control none -k regex:$pat -s 1 -c 1 -f -o /tmp/ncu_${TAG}_$c \
          python scripts/prof_ops.py --reps 2 --only $c > gpurun_out/ncu_${TAG}_$c.log 2>&1
        echo "== ncu $c rc=$?"
        ncu -i /tmp/ncu_${TAG}_$c.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_$c.csv 2>/dev/null
      done
      timeout 600 ncu --set full --clock-control none -k regex:'ddim_step|stft_mel' -c 3 -f -o /tmp/ncu_${TAG}_small \
        python scripts/prof_small.py 2 > gpurun_out/ncu_${TAG}_small.log 2>&1
      echo "== ncu small rc=$?"
      ncu -i /tmp/ncu_${TAG}_small.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_small.csv 2>/dev/null ;;
    micro)        # tcgen05.mma issue rates + per-role timelines of the persistent GEMM (CTA 0)
      timeout 300 python scripts/umma_rate.py > gpurun_out/umma_rate_$TAG.txt 2>&1; cat gpurun_out/umma_rate_$TAG.txt
      timeout 300 python scripts/prof_ops.py --reps 20 --dbg 128 --only ${TL_CASES:-lin_k256_n256,lin_k1024_n256,lin_k256_n2048_geglu,conv_l2_256} > gpurun_out/timeline_$TAG.txt 2>&1
      echo "== timeline"; grep -v "^  " gpurun_out/timeline_$TAG.txt ;;
    step)         # ms per DDIM step of the default configuration
      timeout 300 python scripts/step_time.py --lanes 1 --tag "$TAG" 2>gpurun_out/step_err.log | grep '^{' | tee gpurun_out/step_$TAG.json || tail -3 gpurun_out/step_err.log ;;
    diag)         # what bounds the GEMM: dbg bits 1 skip A loads, 2 skip B loads, 4 skip MMA, 8 skip epilogue; 128 = timeline of CTA 0
      : > gpurun_out/diag_$TAG.txt
      for dbg in ${DIAG_BITS:-0 1 2 3 4 8 12 7 11 128}; do
        echo "## dbg=$dbg" >> gpurun_out/diag_$TAG.txt
        timeout 300 python scripts/prof_ops.py --reps 40 --dbg $dbg --only ${DIAG_CASES:-lin_k256_n256,lin_k1024_n256,lin_k256_n2048_geglu,lin_k640_n640,conv_l1_128,conv_l2_256} >> gpurun_out/diag_$TAG.txt 2>&1
      done
      echo "== diag"; grep -v "^  " gpurun_out/diag_$TAG.txt ;;
    ncu-list)     # per-launch time + DRAM bytes of the UNet: 1 warm-up step (graph capture) + 1 measured DDIM step, graph kernel nodes
      timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none \
        --csv --log-file gpurun_out/launches_$TAG.csv python scripts/step_time.py --steps 1 --warm 1 --reps 1 --tag ncu $NCU_LIST_ARGS > gpurun_out/ncu_list_$TAG.log 2>&1
      echo "== ncu-list rc=$?"; tail -2 gpurun_out/ncu_list_$TAG.log ;;
    *) echo "unknown job $job" ;;
  esac
done
