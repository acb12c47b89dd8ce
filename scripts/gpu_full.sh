#!/bin/bash
# Full round deliverable run on one B200: smoke, gpu tests, bench (both arms), ncu launch list + full capture.
# Usage: gpu_full.sh <tag>
TAG=${1:-r1}
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu_${TAG}.txt; nproc >> $O/gpu_${TAG}.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke_${TAG}.log 2>&1; echo "smoke rc=$?" >> $O/smoke_${TAG}.log
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> $O/pytest_${TAG}.log
timeout 900 python bench.py > $O/bench_${TAG}.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_${TAG}.log 2>&1; echo "rc=$?" >> $O/bench_ref_${TAG}.log
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 300 $BENCH > $O/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_${TAG}.csv $BENCH > $O/ncu_list_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sd_blkr?_apply -s 3 -c 1 -o $O/prof_${TAG} -f $BENCH > $O/ncu_full_${TAG}.log 2>&1
tail -n 2 $O/smoke_${TAG}.log $O/pytest_${TAG}.log $O/bench_ref_${TAG}.log; head -c 2500 $O/bench_${TAG}.log
