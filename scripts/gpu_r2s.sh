#!/bin/bash
# Round 2, call s (1 GPU): the four-configurations-per-lane item body of the single-tail-configuration classes.
TAG=${1:-r2s}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_apply.py tests/test_gpu_solvers.py tests/test_gpu_zzzz_configs.py -q -x 2>&1 | tail -n 5 | tee $O/pytest_${TAG}.txt
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(round(d["ms_per_step"],3), d.get("parity"))
    except Exception: pass'
for st in 5 20; do
  echo "steps=$st: $(timeout 200 python bench.py --steps $st --warmup 3 --no-cpu --no-e2e --no-solve --parity-rows 64 2>&1 | tail -n 1 | python -c "$J")" | tee -a $O/steps_${TAG}.txt
done
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLK_PFP=7"
timeout 300 python bench.py --solve-only --solve-m 30 2>&1 | tail -n 1 | cut -c1-700 | tee $O/solve_${TAG}.txt
