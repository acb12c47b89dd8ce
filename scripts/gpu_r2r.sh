#!/bin/bash
# Round 2, call r (1 GPU): copy-engine test, prefetch default against the number of timed steps, q-batch crossover, periodic.
TAG=${1:-r2r}; O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_copy_engine.py -q -x 2>&1 | tail -n 8 | tee $O/pytest_copy_${TAG}.txt
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(round(d["ms_per_step"],3))
    except Exception: pass'
for st in 5 20 60; do for pf in 0 7; do
  echo "steps=$st SD_BLK_PFP=$pf: $(SD_BLK_PFP=$pf timeout 200 python bench.py --steps $st --warmup 3 --no-cpu --no-e2e --no-solve --no-parity 2>&1 | tail -n 1 | python -c "$J") ms" | tee -a $O/steps_${TAG}.txt
done; done
timeout 300 python scripts/qbatch_crossover.py 16 20 22 24 2>&1 | tail -n 6 | tee $O/qbatch_${TAG}.txt
for L in 28 32; do
  echo "periodic L=$L: $(timeout 300 python bench.py --L $L --boundary periodic --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -n 1 | python -c "$J") ms" | tee -a $O/periodic_${TAG}.txt
done
