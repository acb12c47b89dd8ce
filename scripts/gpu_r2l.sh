#!/bin/bash
# Round 2, call l (N GPUs): sharded parity and the bench lines (L = 32, 34, 36 when they fit) with the default weighted
# shards, plus equal shards for comparison.  Usage: gpurun --gpus N -- 'bash scripts/gpu_r2l.sh <tag> N "32 34"'
TAG=${1:-r2l}; N=${2:-4}; LS=${3:-"32 34"}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
J='import sys,json
for l in sys.stdin:
    try:
        d=json.loads(l); print("ms/apply", round(d["ms_per_step"],3), "parity", d.get("parity"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    except Exception: pass'
timeout 300 $TR scripts/mgpu_check.py > $O/mgpu_${TAG}_n${N}.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}_n${N}.log
echo "== mgpu_check N=$N: $(grep -h 'FAIL\|ALL OK\|rc=' $O/mgpu_${TAG}_n${N}.log | tr '\n' ' ')" | tee -a $O/r2l_${TAG}.txt
for L in $LS; do
  timeout 400 $TR bench.py --gpus $N --L $L --steps 10 --warmup 3 --no-cpu --no-solve > $O/bench_${TAG}_n${N}_L$L.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}_n${N}_L$L.log
  echo "N=$N L=$L weighted shards: $(tail -n 2 $O/bench_${TAG}_n${N}_L$L.log | python -c "$J")" | tee -a $O/r2l_${TAG}.txt
  grep "^{" $O/bench_${TAG}_n${N}_L$L.log | tail -n 1 > $O/benchline_${TAG}_n${N}_L$L.json
done
SD_SHARD_BALANCE=0 timeout 300 $TR bench.py --gpus $N --L 32 --steps 10 --warmup 3 --no-cpu --no-solve --no-e2e > $O/bench_${TAG}_n${N}_L32_equal.log 2>&1
echo "N=$N L=32 equal shards: $(tail -n 2 $O/bench_${TAG}_n${N}_L32_equal.log | python -c "$J")" | tee -a $O/r2l_${TAG}.txt
for c in 100 200; do
  SD_SHARD_REMOTE_COST=$c timeout 300 $TR bench.py --gpus $N --L 32 --steps 10 --warmup 3 --no-cpu --no-solve --no-e2e > $O/bench_${TAG}_n${N}_L32_c$c.log 2>&1
  echo "N=$N L=32 remote cost $c%: $(tail -n 2 $O/bench_${TAG}_n${N}_L32_c$c.log | python -c "$J")" | tee -a $O/r2l_${TAG}.txt
done
