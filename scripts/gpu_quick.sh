#!/bin/bash
# quick check: parity tests, L=32 / L=28 bench, phase timing.  Usage: gpu_quick.sh <tag>
mkdir -p gpurun_out
OUT=gpurun_out/quick_${1:-q}.txt; : > $OUT
timeout 300 python -m pytest tests/test_gpu_apply.py -m gpu -x -q 2>&1 | tail -2 >> $OUT
for L in 32 28; do
  r=$(timeout 200 python bench.py --L $L --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1)
  echo "L=$L -> $r" >> $OUT
done
r=$(timeout 200 python bench.py --L 32 --dtype c128 --steps 5 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'])" 2>&1)
echo "L=32 c128 -> $r" >> $OUT
timeout 200 python scripts/phase_timing.py 32 >> $OUT 2>&1
cat $OUT
