#!/bin/bash
# Round 2, call n (1 GPU): q-batched recurrences + fused reorthogonalisation (new tests first, then the whole gpu suite),
# config 1 / 3 timing, and the A/B of the two apply-kernel experiments (producer L2 bulk prefetch, deeper stream stages).
TAG=${1:-r2n}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py -q -x 2>&1 | tail -n 25 | tee $O/pytest_batch_${TAG}.txt
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -n 15 | tee $O/pytest_${TAG}.txt
timeout 300 python bench.py --configs-only 2>&1 | tail -n 1 | cut -c1-2500 | tee $O/configs_${TAG}.txt
SD_REORTH_FUSED=0 timeout 300 python bench.py --configs-only 2>&1 | tail -n 1 | cut -c1-900 | tee $O/configs_nofuse_${TAG}.txt
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLK_PFP=0" "SD_BLK_PFP=1" "SD_BLK_PFP=3" "SD_BLK_PFP=5" "SD_BLK_PFP=7" "SD_BLKL_DEPTH=3" "SD_BLKL_DEPTH=4" "SD_BLK_PFP=1 SD_BLKL_THREADS=768" "SD_BLK_PFP=5 SD_BLKL_DEPTH=3"
