#!/bin/bash
# Round 2, call c: lean kernel after the LSU diet (constant-bank coupling tables, item-ordered dmid, merged header
# entries, unrolled mid hops): parity, A/B of the CTA size, ncu --set full of the default.
TAG=${1:-r2c}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 | tee $O/pytest_${TAG}.txt
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLKL_THREADS=768" "SD_BLKL_THREADS=640" "SD_BLKL_THREADS=512"
timeout 300 python bench.py --dtype c128 --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | cut -c1-300 | tee $O/bench_c128_${TAG}.txt
timeout 300 python bench.py --L 28 --steps 20 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | cut -c1-300 | tee $O/bench_L28_${TAG}.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:sd_blkl_apply -s 1 -c 1 -f -o $O/prof_${TAG} python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-solve > $O/ncu_${TAG}.log 2>&1
ls -la $O/prof_${TAG}.ncu-rep
