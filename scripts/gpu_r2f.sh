#!/bin/bash
# Round 2, call f (1 GPU): vectorised BLAS-1 + fused Lanczos engine: parity, the full default bench (solve leg), launch
# list of the solve leg, BLAS-1 bandwidth microbench through the ABI.
TAG=${1:-r2f}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8 | tee $O/pytest_${TAG}.txt
timeout 400 python bench.py --no-cpu > $O/bench_${TAG}.log 2>&1; tail -n 1 $O/bench_${TAG}.log | cut -c1-3000 | tee $O/benchline_${TAG}.txt
timeout 300 python scripts/blas_bw.py 2>&1 | tee $O/blas_${TAG}.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_solve_${TAG}.csv python bench.py --solve-only --solve-m 10 > $O/ncu_solve_${TAG}.log 2>&1
tail -n 3 $O/ncu_solve_${TAG}.log
