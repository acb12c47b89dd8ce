#!/bin/bash
# Round 2, call u (1 GPU): sync-free fused reorthogonalisation solve.
TAG=${1:-r2u}; O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_solvers.py tests/test_gpu_sqw_tolerance.py tests/test_gpu_zzz_lean.py -q -x 2>&1 | tail -n 5 | tee $O/pytest_${TAG}.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -n 2 | tee $O/smoke_${TAG}.txt
for i in 1 2; do timeout 300 python bench.py --configs-only 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read())['configs']['config1']; print({k: d[k] for k in ('gpu_s','gpu_launches','gpu_split','E0_abs_diff','Sqw_rel_l2_diff')})" | tee -a $O/config1_${TAG}.txt; done
