#!/bin/bash
# Round 2, call e (8 GPUs): sharded parity at world 8, L=32/34/36 H.psi with the sampled-row parity in the bench line,
# config 5 (L=36 c128 Chebyshev evolution), remote-weighted shards A/B.  Every step under its own timeout.
TAG=${1:-r2e}; N=${2:-8}; O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-solve"
timeout 300 $TR scripts/mgpu_check.py > $O/mgpu_${TAG}_n${N}.log 2>&1; echo "rc=$?" >> $O/mgpu_${TAG}_n${N}.log
echo "== mgpu_check N=$N: $(grep -h 'FAIL\|ALL OK\|rc=' $O/mgpu_${TAG}_n${N}.log | tr '\n' ' ')" | tee -a $O/r2e_${TAG}.txt
for L in 32 34 36; do
  timeout 300 $TR $B --L $L > $O/bench_${TAG}_n${N}_L${L}.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}_n${N}_L${L}.log
  echo "== L=$L N=$N: $(tail -n 2 $O/bench_${TAG}_n${N}_L${L}.log | python -c "import sys,json
for l in sys.stdin:
    try: d=json.loads(l); print('ms/apply', round(d['ms_per_step'],3), 'parity', d.get('parity'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
    except Exception: print(l.strip()[:200])")" | tee -a $O/r2e_${TAG}.txt
done
SD_SHARD_BALANCE=1 SD_SHARD_REMOTE_COST=130 timeout 300 $TR $B --L 32 --no-e2e > $O/bench_${TAG}_n${N}_L32_bal.log 2>&1; echo "rc=$?" >> $O/bench_${TAG}_n${N}_L32_bal.log
echo "== L=32 N=$N balanced shards: $(tail -n 2 $O/bench_${TAG}_n${N}_L32_bal.log | cut -c1-400)" | tee -a $O/r2e_${TAG}.txt
timeout 400 $TR scripts/config5.py > $O/config5_${TAG}_n${N}.log 2>&1; echo "rc=$?" >> $O/config5_${TAG}_n${N}.log
tail -n 2 $O/config5_${TAG}_n${N}.log | cut -c1-900 | tee -a $O/r2e_${TAG}.txt
