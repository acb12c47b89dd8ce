#!/bin/bash
# Round 2, call g (1 GPU): epilogue parameters from the constant bank, native block-layout szq / observables.
TAG=${1:-r2g}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8 | tee $O/pytest_${TAG}.txt
timeout 400 python bench.py --no-cpu --no-e2e > $O/bench_${TAG}.log 2>&1; tail -n 1 $O/bench_${TAG}.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('apply ms', d['ms_per_step'], 'solve', d['solve'])" | tee $O/benchline_${TAG}.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_solve_${TAG}.csv python bench.py --solve-only --solve-m 10 > $O/ncu_solve_${TAG}.log 2>&1
timeout 300 python bench.py --dtype c128 --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c128 apply ms', d['ms_per_step'], d['parity'])" | tee -a $O/benchline_${TAG}.txt
