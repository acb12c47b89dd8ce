#!/bin/bash
# bench (plain) then ncu launch list + one full capture of the apply kernel. Usage: gpu_prof.sh <tag> [extra bench args]
TAG=${1:-r1}; shift
mkdir -p gpurun_out
BENCH="python bench.py --steps 5 --warmup 2 --no-cpu --no-e2e $@"
timeout 300 python -m pytest tests/test_gpu_apply.py -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/bench_${TAG}.log 2>&1
timeout 300 python bench.py --L 28 --steps 20 --warmup 3 --no-cpu --no-e2e "$@" > gpurun_out/bench_${TAG}_L28.log 2>&1
timeout 300 $BENCH > gpurun_out/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${TAG}.csv $BENCH > gpurun_out/ncu_list_${TAG}.log 2>&1
timeout 300 $BENCH > gpurun_out/plain2_${TAG}.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sd_tile_apply -s 2 -c 1 -o gpurun_out/prof_${TAG} -f $BENCH > gpurun_out/ncu_full_${TAG}.log 2>&1
cat gpurun_out/bench_${TAG}.log | head -c 600; echo; tail -n 3 gpurun_out/pytest_${TAG}.log gpurun_out/ncu_full_${TAG}.log
