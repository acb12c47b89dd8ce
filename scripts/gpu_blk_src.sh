#!/bin/bash
# full ncu capture of the block kernel with source import. Usage: gpu_blk_src.sh <tag> [env...]
TAG=$1; shift; O=gpurun_out; mkdir -p $O
P="python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e"
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:sd_blkr?_apply -s 1 -c 1 -o $O/prof_${TAG} -f $P > $O/ncu_full_${TAG}.log 2>&1
tail -2 $O/ncu_full_${TAG}.log
