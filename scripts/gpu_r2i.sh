#!/bin/bash
# Round 2, call i (1 GPU): reduction-before-stores, work-vector pool, fixed S(q,w) test: full GPU suite + default bench.
TAG=${1:-r2i}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 | tee $O/pytest_${TAG}.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee $O/smoke_${TAG}.txt
timeout 900 python bench.py > $O/bench_${TAG}.log 2>&1; tail -n 1 $O/bench_${TAG}.log | cut -c1-8000 > $O/benchline_${TAG}.txt; cut -c1-200 $O/benchline_${TAG}.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_solve_${TAG}.csv python bench.py --solve-only --solve-m 10 > $O/ncu_solve_${TAG}.log 2>&1
