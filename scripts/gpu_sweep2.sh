#!/bin/bash
mkdir -p gpurun_out
OUT=gpurun_out/sweep_${1:-s}.txt; : > $OUT
timeout 300 python -m pytest tests/test_gpu_apply.py -m gpu -x -q 2>&1 | tail -2 >> $OUT
for pf in 0 32 128 512 2048; do
  r=$(SD_PF_DIST=$pf timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1)
  echo "pf=$pf -> $r" >> $OUT
done
for pf in 0 128; do SD_PF_DIST=$pf timeout 200 python scripts/phase_timing.py 32 >> $OUT 2>&1; done
cat $OUT
