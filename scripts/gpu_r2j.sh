#!/bin/bash
# Round 2, call j (1 GPU): producer-side flush of the fused reductions: full GPU suite, solve leg, launch list, ncu full
# of the fused apply, KPM config timing.
TAG=${1:-r2j}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 6 | tee $O/pytest_${TAG}.txt
timeout 300 python bench.py --solve-only --solve-m 30 2>&1 | tail -n 1 | cut -c1-600 | tee $O/solve_${TAG}.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_solve_${TAG}.csv python bench.py --solve-only --solve-m 10 > $O/ncu_solve_${TAG}.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sd_blkl_apply -s 3 -c 1 -f -o $O/prof_fused_${TAG} python bench.py --solve-only --solve-m 6 > $O/ncu_fused_${TAG}.log 2>&1
timeout 300 python bench.py --configs-only 2>&1 | tail -n 1 | cut -c1-1500 | tee $O/configs_${TAG}.txt
