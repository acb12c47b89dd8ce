#!/bin/bash
# NOTE: an 8-GPU call is charged 8x; every step below carries its own short timeout (a hung collective once
# burned the whole round's budget).  Run with: gpurun --gpus 8 --timeout 900 -- 'bash scripts/gpu_scale8.sh <tag>'
# 8-GPU box: sharded parity + L=32/34/36 timing at N=8, bench.py at N=4 and N=8. Usage: gpu_scale8.sh <tag>
TAG=$1; O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_${TAG}.txt 2>&1
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511"; }
timeout 300 $(tr 8) scripts/mgpu_check.py 32 34 36 > $O/mgpu_${TAG}_8.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}_8.log
for n in 8 4; do
  timeout 240 $(tr $n) bench.py --gpus $n --steps 10 --warmup 3 > $O/bench_${TAG}_$n.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}_$n.log
done
timeout 240 $(tr 4) scripts/mgpu_check.py 32 34 > $O/mgpu_${TAG}_4.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}_4.log
grep -h "FAIL\|ALL OK\|ms/apply\|rc=\|Error\|error" $O/mgpu_${TAG}_8.log $O/mgpu_${TAG}_4.log | tail -n 30
for n in 8 4; do tail -n 2 $O/bench_${TAG}_$n.log | cut -c1-600; done
