#!/bin/bash
# Round 2, call k (1 GPU): code-size diet (mid loop, noinline division, compile-time epilogue kinds), sync-free KPM.
TAG=${1:-r2k}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -n 6 | tee $O/pytest_${TAG}.txt
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLKL_THREADS=640" "SD_BLKL_THREADS=768"
timeout 300 python bench.py --solve-only --solve-m 30 2>&1 | tail -n 1 | cut -c1-600 | tee $O/solve_${TAG}.txt
timeout 300 python bench.py --configs-only 2>&1 | tail -n 1 | cut -c1-1500 | tee $O/configs_${TAG}.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_solve_${TAG}.csv python bench.py --solve-only --solve-m 10 > $O/ncu_solve_${TAG}.log 2>&1
timeout 300 python bench.py --dtype c128 --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c128 apply ms', d['ms_per_step'], d['parity'])" | tee -a $O/solve_${TAG}.txt
