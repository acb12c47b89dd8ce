#!/bin/bash
# Last check of the tree as shipped (1 GPU): smoke, the whole gpu suite, one short bench line.
TAG=${1:-final}; O=gpurun_out; mkdir -p $O
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -n 2 | tee $O/smoke_${TAG}.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -n 4 | tee $O/pytest_${TAG}.txt
timeout 400 python bench.py --steps 20 --warmup 3 --no-solve --no-configs 2>&1 | tail -n 1 | cut -c1-2500 | tee $O/bench_${TAG}.txt
