#!/bin/bash
# Round 2, call d (N GPUs): reordered lean kernel on 1 GPU, then sharded parity + timing with direct peer loads and
# with the halo mirror.  Usage: gpurun --gpus 2 --timeout 900 -- 'bash scripts/gpu_r2d.sh r2d 2'
TAG=${1:-r2d}; N=${2:-2}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for t in 640 512; do
  r=$(SD_BLKL_THREADS=$t timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-solve 2>&1 | tail -1 | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])" 2>&1 | tail -1)
  echo "1 GPU threads=$t -> $r ms" | tee -a $O/r2d_${TAG}.txt
done
for halo in 0 1; do
  env SD_HALO=$halo timeout 300 $TR scripts/mgpu_check.py 32 > $O/mgpu_${TAG}_n${N}_h${halo}.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}_n${N}_h${halo}.log
  echo "== N=$N SD_HALO=$halo: $(grep -h 'FAIL\|ALL OK\|ms/apply\|rc=' $O/mgpu_${TAG}_n${N}_h${halo}.log | tr '\n' ' ')" | tee -a $O/r2d_${TAG}.txt
  env SD_HALO=$halo timeout 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-solve > $O/bench_${TAG}_n${N}_h${halo}.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}_n${N}_h${halo}.log
  tail -n 2 $O/bench_${TAG}_n${N}_h${halo}.log | cut -c1-700 | tee -a $O/r2d_${TAG}.txt
done
