#!/bin/bash
# correctness + decomposition timing of the block kernel. Usage: gpu_blk_quick.sh <tag> [dbg list]
TAG=${1:-q}; shift; O=gpurun_out; mkdir -p $O
timeout 600 python scripts/blk_check.py 28 2>&1 | tail -4 > $O/blkq_${TAG}.txt
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e"
for dbg in ${@:-0 1 2 3 8}; do
  r=$(SD_BLK_DBG=$dbg timeout 200 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])" 2>&1)
  echo "dbg=$dbg -> $r ms" >> $O/blkq_${TAG}.txt
done
cat $O/blkq_${TAG}.txt
