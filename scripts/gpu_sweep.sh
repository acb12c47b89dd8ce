#!/bin/bash
# sweep tile-kernel launch parameters: SD_TILE_THREADS x SD_TILE_B (x SD_FAR_MB)
mkdir -p gpurun_out
OUT=gpurun_out/sweep_${1:-s}.txt; : > $OUT
for cfg in "512 15 100" "256 15 100" "256 14 100" "512 14 100" "256 13 100" "512 13 100" "256 14 0" "256 14 100000" "256 12 100"; do
  set -- $cfg
  r=$(SD_TILE_THREADS=$1 SD_TILE_B=$2 SD_FAR_MB=$3 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1)
  echo "threads=$1 B=$2 far=$3 -> $r" >> $OUT
done
cat $OUT
