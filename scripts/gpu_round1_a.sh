#!/bin/bash
# first GPU contact: smoke, parity tests, a short bench at L=28 and L=32
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py --L 28 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_L28.log 2>&1
timeout 300 python bench.py --L 28 --steps 10 --warmup 3 --no-cpu --no-e2e --path generic > gpurun_out/bench_L28_generic.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_L32.log 2>&1
tail -5 gpurun_out/smoke.log gpurun_out/pytest_gpu.log gpurun_out/bench_L28.log gpurun_out/bench_L28_generic.log gpurun_out/bench_L32.log
