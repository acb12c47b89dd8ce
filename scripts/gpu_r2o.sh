#!/bin/bash
# Round 2, call o (1 GPU): periodic chain on the block path, copy engine / two-deep e2e, prefetch variants, config timings.
TAG=${1:-r2o}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 12 | tee $O/pytest_${TAG}.txt
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLK_PFP=7" "SD_BLK_PFP=23" "SD_BLK_PFP=15" "SD_BLK_PFP=31" "SD_BLK_PFP=23 SD_BLKL_THREADS=768"
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print("ms/apply", round(d["ms_per_step"],3), "launches", d["gpu_launches"], "parity", d.get("parity"), "e2e", d.get("e2e"))
    except Exception: pass'
for L in 28 32; do
  echo "periodic L=$L: $(timeout 300 python bench.py --L $L --boundary periodic --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | python -c "$J")" | tee -a $O/periodic_${TAG}.txt
done
echo "periodic L=28 generic: $(timeout 300 python bench.py --L 28 --boundary periodic --path generic --steps 5 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | python -c "$J")" | tee -a $O/periodic_${TAG}.txt
echo "open L=32 e2e: $(timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu --no-solve 2>&1 | tail -n 1 | python -c "$J")" | tee $O/e2e_${TAG}.txt
for i in 1 2; do timeout 300 python bench.py --configs-only 2>&1 | tail -n 1 | cut -c1-2600 | tee -a $O/configs_${TAG}.txt; done
timeout 300 python bench.py --solve-only --solve-m 30 2>&1 | tail -n 1 | cut -c1-700 | tee $O/solve_${TAG}.txt
