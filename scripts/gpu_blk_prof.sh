#!/bin/bash
# decomposition runs + ncu of the block kernel. Usage: gpu_blk_prof.sh <tag>
TAG=${1:-b1}; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e"
for dbg in 0 1 2 3 4 7; do
  r=$(SD_BLK_DBG=$dbg timeout 200 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])" 2>&1)
  echo "dbg=$dbg -> $r ms" >> $O/blkdbg_${TAG}.txt
done
for nb in 2 4; do
  r=$(SD_BLK_NBUF=$nb timeout 200 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])" 2>&1)
  echo "nbuf=$nb -> $r ms" >> $O/blkdbg_${TAG}.txt
done
for far in 0 100000; do
  r=$(SD_FAR_MB=$far timeout 200 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])" 2>&1)
  echo "far_mb=$far -> $r ms" >> $O/blkdbg_${TAG}.txt
done
cat $O/blkdbg_${TAG}.txt
P="python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sd_blkr?_apply -s 2 -c 1 -o $O/prof_${TAG} -f $P > $O/ncu_full_${TAG}.log 2>&1
tail -2 $O/ncu_full_${TAG}.log
