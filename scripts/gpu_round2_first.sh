#!/bin/bash
# First GPU call of round 2 (one GPU, ~12 min): everything that was written after the round-1 GPU budget ran out.
#   1. the default path still green (pytest -m gpu, smoke, bench)
#   2. ring kernel (sd_blkr.h): parity vs the oracle under a timeout + watchdog, then time at L=28/32
#   3. A/B of every knob on the L=32 bench: time, DRAM bytes, L2 hit rate, instructions
#        SD_BLK_RING=1            ring kernel (TMA-staged neighbour tiles, tile-stationary register accumulators)
#        SD_BLK_ORDER=2 / =1      breadth-first / grouped-greedy tile order (LRU model: -41 % / -20 % DRAM reads)
#        SD_BLK_VARIANT=1         lean item body of the standard kernel
# Usage: gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh r2a'
TAG=${1:-r2a}; O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu_${TAG}.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 5 | tee $O/pytest_${TAG}.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee $O/smoke_${TAG}.txt
timeout 300 python bench.py > $O/bench_${TAG}.log 2>&1; tail -n 1 $O/bench_${TAG}.log
# L2 -> SM throughput of this GPU (DESIGN 4.1a argues from ~6300 B/clk = 12.2 TB/s; measure it)
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/mb_l2cap scripts/mb_l2cap.cu 2> /dev/null && timeout 120 scripts/mb_l2cap 2>&1 | tee $O/l2cap_${TAG}.txt | tail -n 12
echo "== ring kernel parity" | tee $O/ring_${TAG}.txt
# diagnostic watchdog first (SD_BLK_DBG=64: a stuck barrier wait ends the launch with a record of where, instead of a trap)
SD_BLK_RING=1 SD_BLK_DBG=64 timeout 300 python scripts/ring_check.py 2>&1 | tail -n 25 | tee -a $O/ring_${TAG}.txt
SD_BLK_RING=1 timeout 300 python scripts/ring_check.py 28 32 2>&1 | tail -n 8 | tee -a $O/ring_${TAG}.txt
SD_BLK_RING=1 SD_BLK_ORDER=2 timeout 300 python scripts/ring_check.py 32 2>&1 | tail -n 3 | tee -a $O/ring_${TAG}.txt
SD_BLK_RING=1 SD_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_apply.py tests/test_gpu_solvers.py -m gpu -x -q 2>&1 | tail -n 3 | tee -a $O/ring_${TAG}.txt
bash scripts/gpu_blk_ncu2.sh ${TAG} "SD_BLK_RING=0" "SD_BLK_RING=1" "SD_BLK_RING=1 SD_BLK_ORDER=2" "SD_BLK_ORDER=2" "SD_BLK_ORDER=1" \
     "SD_BLK_VARIANT=1" "SD_BLK_ORDER=2 SD_BLK_VARIANT=1" "SD_BLK_RING=1 SD_BLK_DBG=1" "SD_BLK_RING=1 SD_BLK_DBG=3" "SD_BLK_RING=1 SD_BLK_ORDER=2 SD_BLK_ORDER_E=14" "SD_BLK_RING=1 SD_BLK_DBG=16" "SD_BLK_RING=1 SD_BLK_DBG=32" "SD_BLK_RING=1 SD_BLKR_DIRECT=2" "SD_BLK_RING=1 SD_BLKR_DIRECT=4" "SD_BLK_RING=1 SD_BLKR_DIRECT=4 SD_BLK_ORDER=2"
