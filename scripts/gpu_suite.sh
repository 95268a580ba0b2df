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
    bench2) TNN_GEMM_CG=2 timeout 900 python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cg2.log 2>&1 ;;
    mnist) timeout 600 python bench.py --workload mnist --steps 2000 --warmup 50 > gpurun_out/bench_mnist.log 2>&1; timeout 600 python bench.py --workload mnist --steps 2000 --warmup 50 --graph off --no-cpu-baseline > gpurun_out/bench_mnist_eager.log 2>&1 ;;
    sweep) timeout 900 python scripts/sweep_ops.py --cpu > gpurun_out/sweep.json 2> gpurun_out/sweep.err ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_list.log 2>&1 ;;
    ncu_gemm) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16x3 -s 22 -c 3 -f -o gpurun_out/prof_gemm python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_gemm.log 2>&1; ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null; [ $(stat -c %s gpurun_out/prof_gemm.ncu-rep) -gt 20000000 ] && rm -f gpurun_out/prof_gemm.ncu-rep ;;
    dist)  timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_dist.log 2>&1 ;;
    bench_n2) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1 ;;
    bench_n4) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4.log 2>&1 ;;
    bench_n8) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_n8.log 2>&1 ;;
    bench1) timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1 ;;
    memcheck) timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py tests/test_gpu_autograd_cases.py tests/test_gpu_gemm.py tests/test_gpu_graph.py tests/test_gpu_train.py -m gpu -q -x -p no:cacheprovider > gpurun_out/memcheck.log 2>&1 ;;
    gemmbench) timeout 600 python scripts/gemm_bench.py --cg 2 --ksplit 0 --group-m 1,-8,-4,-2,1,-8,-4,-2 > gpurun_out/gemm_bench.jsonl 2>&1 ;;
    ab_fuse) for i in 1 2 3; do TNN_FUSE_RELU=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline >> gpurun_out/ab_fuse1.log 2>&1; TNN_FUSE_RELU=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline >> gpurun_out/ab_fuse0.log 2>&1; done ;;
    ncu_mnist) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_mnist.csv python bench.py --workload mnist --steps 20 --warmup 10 --no-cpu-baseline > gpurun_out/ncu_mnist.log 2>&1 ;;
    ab_bwd) for i in 1 2 3; do TNN_FUSE_RELU_BWD=1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline >> gpurun_out/ab_bwd1.log 2>&1; TNN_FUSE_RELU_BWD=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline >> gpurun_out/ab_bwd0.log 2>&1; done ;;
    ncu_mem) timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:adam_vec|relu_kernel|split_tf32|reduce_col|ce_bwd|ce_rows|ce_partial' -s 30 -c 18 -f -o gpurun_out/prof_mem python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_mem.log 2>&1; ncu -i gpurun_out/prof_mem.ncu-rep --page raw --csv > gpurun_out/prof_mem_raw.csv 2>/dev/null; [ $(stat -c %s gpurun_out/prof_mem.ncu-rep) -gt 20000000 ] && rm -f gpurun_out/prof_mem.ncu-rep ;;
    gemmbench2) timeout 600 python scripts/gemm_bench.py --cg 2 --ksplit 0 --group-m 1 --sustain 500 > gpurun_out/gemm_bench_mix.jsonl 2>&1; TNN_GEMM_SPLIT=tf32x3 timeout 600 python scripts/gemm_bench.py --cg 2 --ksplit 0 --group-m 1 --sustain 500 > gpurun_out/gemm_bench_tf32x3.jsonl 2>&1 ;;
    graph) timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_graph.log 2>&1 ;;
    mnist_ab) for g in on off on off; do timeout 300 python bench.py --workload mnist --steps 2000 --warmup 50 --graph $g --no-cpu-baseline >> gpurun_out/bench_mnist_graph_$g.log 2>&1; done ;;
    ncu_mnist_graph) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -s 100 -c 40 --csv --log-file gpurun_out/launches_mnist_graph.csv python bench.py --workload mnist --steps 20 --warmup 10 --graph on --no-cpu-baseline > gpurun_out/ncu_mnist_graph.log 2>&1 ;;
    diag)  timeout 600 python scripts/diag_fp32_traj.py > gpurun_out/diag_fp32.jsonl 2>&1 ;;
    all)   timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1 ;;
  esac
  echo "stage $s exit $?" >> gpurun_out/stages.txt
done
for f in gpurun_out/diag_fp32.jsonl gpurun_out/smoke.log gpurun_out/gemm_bench_mix.jsonl gpurun_out/gemm_bench_tf32x3.jsonl gpurun_out/pytest_graph.log gpurun_out/pytest_all.log gpurun_out/bench_mnist_graph_on.log gpurun_out/bench_mnist_graph_off.log gpurun_out/ncu_mnist_graph.log gpurun_out/pytest_ops.log gpurun_out/pytest_train.log gpurun_out/pytest_gemm.log gpurun_out/pytest_large.log gpurun_out/bench.log gpurun_out/bench_cg2.log gpurun_out/bench_mnist.log gpurun_out/sweep.err gpurun_out/ncu_list.log gpurun_out/ncu_gemm.log gpurun_out/memcheck.log gpurun_out/pytest_dist.log gpurun_out/bench_n1.log gpurun_out/bench_n2.log gpurun_out/bench_n4.log gpurun_out/bench_n8.log; do
  [ -f $f ] && { echo "== $f"; tail -n 6 $f; }
done
cat gpurun_out/stages.txt
du -sh gpurun_out | tail -1
