#!/bin/bash
# Round 2, call v (1 GPU): add-in as a compile-time kernel variant -- open chain timing, periodic parity.
TAG=${1:-r2v}; O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_apply.py -q -x 2>&1 | tail -n 2 | tee $O/pytest_${TAG}.txt
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print(round(d["ms_per_step"],3), (d.get("parity") or {}).get("max_rel_err"))
    except Exception: pass'
for st in 5 20; do
  echo "open steps=$st: $(timeout 200 python bench.py --steps $st --warmup 3 --no-cpu --no-e2e --no-solve --parity-rows 32 2>&1 | tail -n 1 | python -c "$J")" | tee -a $O/steps_${TAG}.txt
done
echo "periodic L=32: $(timeout 200 python bench.py --boundary periodic --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve --parity-rows 32 2>&1 | tail -n 1 | python -c "$J")" | tee -a $O/steps_${TAG}.txt
