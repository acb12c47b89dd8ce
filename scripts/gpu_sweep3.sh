#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/sweep_${1:-s}.txt; : > $OUT
timeout 300 python -m pytest tests/test_gpu_apply.py -m gpu -x -q 2>&1 | tail -2 >> $OUT
for cfg in "512 15 128" "512 15 0" "256 15 128" "256 14 128" "512 14 128"; do
  set -- $cfg
  r=$(SD_TILE_THREADS=$1 SD_TILE_B=$2 SD_PF_DIST=$3 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1)
  echo "threads=$1 B=$2 pf=$3 -> $r" >> $OUT
done
SD_PF_DIST=128 timeout 200 python scripts/phase_timing.py 32 >> $OUT 2>&1
SD_TILE_THREADS=256 SD_TILE_B=14 SD_PF_DIST=128 timeout 200 python scripts/phase_timing.py 32 >> $OUT 2>&1
cat $OUT
