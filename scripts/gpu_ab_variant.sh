#!/bin/bash
# A/B of the block kernel's item-body variants on one GPU: parity (blk_check + pytest subset), time, DRAM bytes,
# instruction count.  Usage: gpu_ab_variant.sh <tag>
TAG=${1:-ab}; O=gpurun_out; mkdir -p $O
for v in 0 1; do
  echo "== SD_BLK_VARIANT=$v" | tee -a $O/ab_${TAG}.txt
  SD_BLK_VARIANT=$v timeout 300 python scripts/blk_check.py 28 32 2>&1 | tail -n 4 | tee -a $O/ab_${TAG}.txt
  SD_BLK_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_apply.py tests/test_gpu_solvers.py -m gpu -x -q 2>&1 | tail -n 2 | tee -a $O/ab_${TAG}.txt
done
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLK_VARIANT=0" "SD_BLK_VARIANT=1" "SD_BLK_ORDER=2" "SD_BLK_ORDER=2 SD_BLK_ORDER_E=14" "SD_BLK_ORDER=1" "SD_BLK_ORDER=2 SD_BLK_VARIANT=1" "SD_BLK_VARIANT=0 SD_BLK_DBG=2" "SD_BLK_VARIANT=1 SD_BLK_DBG=2"
for e in "SD_BLK_ORDER=2" "SD_BLK_ORDER=1" "SD_BLK_ORDER=2 SD_BLK_VARIANT=1"; do env $e timeout 300 python scripts/blk_check.py 2>&1 | tail -n 1 | tee -a $O/ab_${TAG}.txt; done
