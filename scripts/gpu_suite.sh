#!/bin/bash
# Runs the GPU checks on a B200 box (via gpurun), one process per stage so a sticky CUDA error in
# one stage cannot poison the next; logs go to gpurun_out/.
#   scripts/gpu_suite.sh [stages...]   stages: smoke ops train gemm large bench ncu
set -u
mkdir -p gpurun_out
STAGES=${@:-"smoke ops train gemm large bench"}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
for s in $STAGES; do
  case $s in
    smoke) timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ;;
    ops)   timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_autograd_cases.py -m gpu -q -p no:cacheprovider --maxfail=40 > gpurun_out/pytest_ops.log 2>&1 ;;
    train) timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1 ;;
    gemm)  timeout 900 python -m pytest tests/test_gpu_gemm.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1 ;;
    large) timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_large.log 2>&1 ;;
    bench) timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2>&1 ;;
    all)   timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1 ;;
  esac
  echo "stage $s exit $?" >> gpurun_out/stages.txt
done
for f in gpurun_out/smoke.log gpurun_out/pytest_ops.log gpurun_out/pytest_train.log gpurun_out/pytest_gemm.log gpurun_out/pytest_large.log gpurun_out/bench.log; do
  [ -f $f ] && { echo "== $f"; tail -n 6 $f; }
done
cat gpurun_out/stages.txt
