#!/bin/bash
# Round 2, call h (1 GPU): new GPU tests (site-resolved KPM, S(q,w) tolerance), A/B of the per-item L2 prefetch and of
# the reduction fences, full default bench line (solve + config legs).
TAG=${1:-r2h}; O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 8 | tee $O/pytest_${TAG}.txt
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLK_PREFETCH=0" "SD_BLK_PREFETCH=2" "SD_BLK_PREFETCH=1" "SD_BLK_PREFETCH=4"
for nf in 0 1; do
  echo "SD_BLK_NOFENCE=$nf: $(SD_BLK_NOFENCE=$nf timeout 300 python bench.py --solve-only --solve-m 20 2>&1 | tail -n 1 | cut -c1-400)" | tee -a $O/fence_${TAG}.txt
done
timeout 600 python bench.py > $O/bench_${TAG}.log 2>&1; tail -n 1 $O/bench_${TAG}.log | cut -c1-6000 > $O/benchline_${TAG}.txt; cut -c1-300 $O/benchline_${TAG}.txt
