#!/bin/bash
# Round 2, call b: the lean block kernel (sd_blkl.h) on the GPU for the first time.
#   parity (pytest -m gpu), A/B on the L=32 bench (CTA size, tile order, round-1 body), ncu --set full of the default.
TAG=${1:-r2b}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 6 | tee $O/pytest_${TAG}.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee $O/smoke_${TAG}.txt
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLK_KERNEL=1" "SD_BLKL_THREADS=512" "SD_BLKL_THREADS=768" "SD_BLK_ORDER=0" "SD_BLK_KERNEL=0" "SD_BLK_KERNEL=0 SD_BLK_ORDER=0"
timeout 300 python bench.py --dtype c128 --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | cut -c1-300 | tee $O/bench_c128_${TAG}.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sd_blkl_apply -s 1 -c 1 -f -o $O/prof_${TAG} python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-solve > $O/ncu_${TAG}.log 2>&1
ls -la $O/prof_${TAG}.ncu-rep
