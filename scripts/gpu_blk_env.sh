#!/bin/bash
# time the block kernel under env settings. Usage: gpu_blk_env.sh <tag> "ENV1=a ENV2=b" "ENV3=c" ...
TAG=$1; shift; O=gpurun_out; mkdir -p $O
B="python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e"
for e in "$@"; do
  r=$(env $e timeout 200 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])" 2>&1 | tail -1)
  echo "$e -> $r ms" | tee -a $O/blkenv_${TAG}.txt
done
