#!/bin/bash
# Round 2, call q (1 GPU): final evidence pass of the shipped default -- smoke, gpu tests, the full bench line, the reference
# arm at L = 32, periodic L = 32, ncu launch list of the bench command and one ncu --set full capture of the apply kernel.
TAG=${1:-r2q}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/gpu_${TAG}.txt; nproc >> $O/gpu_${TAG}.txt; free -g | head -2 >> $O/gpu_${TAG}.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $O/smoke_${TAG}.log 2>&1; echo "smoke rc=$?" >> $O/smoke_${TAG}.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> $O/pytest_${TAG}.log
timeout 900 python bench.py > $O/bench_${TAG}.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}.log
BENCH="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-solve --no-parity"
timeout 300 $BENCH > $O/plain_${TAG}.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/launches_${TAG}.csv $BENCH > $O/ncu_list_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sd_blkl_apply -s 3 -c 1 -o $O/prof_${TAG} -f $BENCH > $O/ncu_full_${TAG}.log 2>&1
timeout 300 python bench.py --L 32 --boundary periodic --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | cut -c1-1200 > $O/periodic_${TAG}.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_${TAG}.log 2>&1; echo "rc=$?" >> $O/bench_ref_${TAG}.log
tail -n 3 $O/smoke_${TAG}.log $O/pytest_${TAG}.log; tail -c 1500 $O/bench_ref_${TAG}.log; head -c 3000 $O/bench_${TAG}.log
