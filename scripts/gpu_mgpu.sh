#!/bin/bash
# sharded parity + timing + bench on N GPUs of one box. Usage: gpu_mgpu.sh <tag> <N> [L_time ...]
TAG=$1; N=$2; shift; shift
O=gpurun_out; mkdir -p $O
nvidia-smi topo -m > $O/topo_${TAG}.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/mgpu_check.py "$@" > $O/mgpu_${TAG}.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}.log
timeout 240 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${TAG}.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}.log
grep -v "^\[W\|^W1\|Warning" $O/mgpu_${TAG}.log | tail -n 60; tail -n 3 $O/bench_${TAG}.log | cut -c1-1500
